timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "nx512 or mu_512" 2>&1 | tail -4 || exit 1
timeout 120 python tools/bench_variants.py --quick --size 512 > gpurun_out/r02_variants_512.json 2> gpurun_out/x16_1.err; cat gpurun_out/r02_variants_512.json; tail -2 gpurun_out/x16_1.err
