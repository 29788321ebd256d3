#!/bin/bash
# Experiment build (NOT the product): librnnt_b200 with the tensor-core products of joint_cg_mm.cu compiled out
# (-DRNNTB200_EXP_NO_MMA), next to the product objects.  Used by scripts/exp_mma_share.py to measure how much of
# cg_lse_mm / cg_grad_mm is the product itself -- the part a tcgen05 formulation could speed up.
set -e
cd "$(dirname "$0")/../rnntransducer_b200"
python -m rnntransducer_b200.build >/dev/null 2>&1 || (cd .. && python -m rnntransducer_b200.build >/dev/null)
mkdir -p build/exp
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     -DRNNTB200_EXP_NO_MMA -c -o build/exp/joint_cg_mm.o csrc/joint_cg_mm.cu
objs=$(ls build/*.o | grep -v joint_cg_mm.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/exp/librnnt_b200_nomma.so $objs build/exp/joint_cg_mm.o -lcuda 2>/dev/null || \
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/exp/librnnt_b200_nomma.so $objs build/exp/joint_cg_mm.o
ls -la build/exp/librnnt_b200_nomma.so
