#!/bin/bash
# GPU box, one GPU: ncu --set full of the k_ao launch of one S2 frame, shipped build (left child first) and the A/B build
# with the reference's near-child-first order (GB_ANY_FIXED_ORDER=0).
out=gpurun_out; mkdir -p $out
timeout 120 python tools/one_render.py bunny_ao || exit 1
timeout 200 ncu --set full --import-source on --clock-control none -k k_ao -c 1 -f -o $out/r02m_k_ao_left_first python tools/one_render.py bunny_ao > $out/ncu22a.log 2>&1
GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_nearfirst.so timeout 200 ncu --set full --import-source on --clock-control none -k k_ao -c 1 -f -o $out/r02m_k_ao_near_first python tools/one_render.py bunny_ao > $out/ncu22b.log 2>&1
ls -la $out/r02m_* | cut -c1-120
