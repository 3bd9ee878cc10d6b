timeout 1200 python -m pytest tests -m gpu -q -x -k "not 512_against_oracle" 2>&1 | tail -8
for S in 1 0; do for n in 512 1024; do T21_X_SEGMENTS=$S timeout 600 python tools/bench_variants.py --quick --size $n > gpurun_out/xseg_${S}_$n.json 2> gpurun_out/xseg_${S}_$n.err; echo "XSEG=$S n=$n"; cat gpurun_out/xseg_${S}_$n.json; tail -2 gpurun_out/xseg_${S}_$n.err; done; done
