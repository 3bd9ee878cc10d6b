for n in 250 300 600; do timeout 300 python tools/bench_variants.py --size $n --quick > gpurun_out/r02_variants_$n.json 2> gpurun_out/r02_variants_$n.err; echo "rc=$? $n"; done
