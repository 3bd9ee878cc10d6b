timeout 1200 python -m pytest tests -m gpu -q -x -k "not 512_against_oracle" 2>&1 | tail -6
timeout 600 python tools/bench_variants.py --quick > gpurun_out/ovf_variants.json 2> gpurun_out/ovf_variants.err; cat gpurun_out/ovf_variants.json; tail -3 gpurun_out/ovf_variants.err
