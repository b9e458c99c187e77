"""Loader for the committed reference fixtures (tests/golden/*.npz, written by oracle/make_goldens.py)."""
import glob
import json
import os
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    g = SimpleNamespace(name=name, meta=meta, model_conf=meta["model_conf"], renderer_conf=meta["renderer_conf"],
                        ids_encoder=meta["ids_encoder"], ids_render=meta["ids_render"], sd={}, out={}, query={}, grad={})
    for k in z.files:
        if k == "meta":
            continue
        t = torch.from_numpy(z[k])
        if k.startswith("sd/"):
            g.sd[k[3:]] = t
        elif k.startswith("out/"):
            g.out[k[4:]] = t
        elif k.startswith("query/"):
            g.query[k[6:]] = t
        elif k.startswith("grad/"):
            g.grad[k[5:]] = t
        else:
            setattr(g, k, t)
    return g


def mlp_params(sd):
    """state-dict entries of mlp_coarse with the prefix stripped."""
    return {k[len("mlp_coarse."):]: v for k, v in sd.items() if k.startswith("mlp_coarse.")}
