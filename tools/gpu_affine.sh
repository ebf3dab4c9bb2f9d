mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_affine.py -m gpu -x -q > gpurun_out/r05h_affine_pytest.log 2>&1; tail -4 gpurun_out/r05h_affine_pytest.log
timeout 30 python tools/affine_bench.py --gaps 20 --check 32 > gpurun_out/r05h_affine_bench.json 2> gpurun_out/r05h_affine_bench.err; cut -c1-420 gpurun_out/r05h_affine_bench.json
timeout 50 python bench.py --config affine --cpu-budget 4 > gpurun_out/r05h_bench_affine.json 2> gpurun_out/r05h_bench_affine.err; tail -c 600 gpurun_out/r05h_bench_affine.json
