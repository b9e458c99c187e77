#!/bin/bash
# builds tools/bin/umma_probe (tcgen05 operand-path probe; see tools/umma_probe.cu)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -lineinfo -std=c++17 -o tools/bin/umma_probe tools/umma_probe.cu
