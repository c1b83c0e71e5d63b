#!/bin/bash
# round-end evidence on one B200: tests, bench lines, ncu launch list + full-set capture
python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 600 gpurun_out/r02_bench_c3.json
for wl in c2 c4 c5; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c3_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
SGPR_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"desc_forward|desc_backward|i8gemm|neighbor_bin" -s 12 -c 6 -o gpurun_out/r02_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_full.ncu-rep
