#!/bin/bash
# The first GPU call of a round, in one box: (1) the GPU suite on the library default (wide BVH + kd arbiter), (2) the same
# suite on the kd-only traversal (RGK_TEST_TRAVERSAL=kd, read by tests/conftest.py), (3) the rgk_device_cfg sweep of the
# headline round, (4) the kd / BVH A/B with full-size bit comparison.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/suite_bvh.log 2>&1; echo "rc=$?" >> gpurun_out/suite_bvh.log
RGK_TEST_TRAVERSAL=kd timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/suite_kd.log 2>&1; echo "rc=$?" >> gpurun_out/suite_kd.log
timeout 120 python tools/bvh_ab.py --sweep --render 2 > gpurun_out/bvh_sweep.json 2> gpurun_out/bvh_sweep.err
timeout 120 python tools/bvh_ab.py > gpurun_out/bvh_ab.json 2> gpurun_out/bvh_ab.err
tail -n 3 gpurun_out/suite_kd.log gpurun_out/suite_bvh.log
