timeout 900 python -m pytest tests -m gpu -q -x -k "not 512_against_oracle" 2>&1 | tail -8
timeout 600 python tools/bench_variants.py --quick > gpurun_out/seg_variants2.json 2> gpurun_out/seg_variants2.err; cat gpurun_out/seg_variants2.json; tail -3 gpurun_out/seg_variants2.err
