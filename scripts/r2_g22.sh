mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_broadphase.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
bash scripts/r2_ab.sh
