timeout 900 python -m pytest tests -m gpu -q -x -k "mixed_and_odd or c2ray or golden or degenerate or shapes or inputs" 2>&1 | tail -6
for n in 250 300 600; do timeout 300 python tools/bench_variants.py --size $n --quick > gpurun_out/r02_variants_$n.json 2>gpurun_out/r02_variants_$n.err; echo "$n"; cat gpurun_out/r02_variants_$n.json; tail -2 gpurun_out/r02_variants_$n.err; done
bash tools/r2_np2b.sh
