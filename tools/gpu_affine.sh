mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_affine.py -m gpu -x -q > gpurun_out/r05e_affine_pytest.log 2>&1; tail -15 gpurun_out/r05e_affine_pytest.log
timeout 100 python tools/affine_bench.py --gaps 20 > gpurun_out/r05e_affine_bench.json 2> gpurun_out/r05e_affine_bench.err; tail -c 1500 gpurun_out/r05e_affine_bench.json; tail -5 gpurun_out/r05e_affine_bench.err
