#!/bin/bash
# The first GPU call of a round, in one box: (1) the GPU suite as shipped, (2) the same suite with every scene committed under
# RGK_WIDE_BVH=1 (the gate for making the wide-BVH path the library default), (3) the knob sweep of the BVH round (includes the
# A/B candidates RGK_SHADE_LAST and RGK_BVH_SHADOW_NOSORT), (4) the kd / BVH A/B with full-size bit comparison.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/suite_kd.log 2>&1; echo "rc=$?" >> gpurun_out/suite_kd.log
RGK_TEST_TRAVERSAL=bvh timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/suite_bvh.log 2>&1; echo "rc=$?" >> gpurun_out/suite_bvh.log
timeout 120 python tools/bvh_ab.py --sweep --render 2 > gpurun_out/bvh_sweep.json 2> gpurun_out/bvh_sweep.err
timeout 120 python tools/bvh_ab.py > gpurun_out/bvh_ab.json 2> gpurun_out/bvh_ab.err
tail -3 gpurun_out/suite_kd.log gpurun_out/suite_bvh.log
