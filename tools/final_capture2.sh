#!/bin/bash
# Short re-validation after a kernel change: parity tests, variant ranking, the bench line, one ncu capture.
mkdir -p gpurun_out
(time timeout 200 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log
(cd tools/variants && ./out/variant_bench 234 out/*.cubin) > gpurun_out/variants_final.txt 2>&1
cat gpurun_out/variants_final.txt
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-260 gpurun_out/bench_final.json
ncu --set full --clock-control none --import-source on -k regex:"k_leaf_hash" -s 1 -c 1 \
    -o gpurun_out/prof_leaf_final python bench.py --steps 1 --warmup 1 --log-n 16 --no-cpu --no-ntt --no-extra > gpurun_out/prof_leaf_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-ntt --no-extra > gpurun_out/launches_bench.log 2>&1
ls gpurun_out | tail -3
