#!/bin/bash
# end-of-round evidence: GPU tests, the bench line, the ncu launch list and one full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
gzip -f gpurun_out/launches.csv
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:krylov_cgs2|slu_fused|slu_bwd_stage|slu_fwd_stage|bell_matvec" -s 600 -c 10 -f -o gpurun_out/full_r1 python scripts/quick_bench.py 10001 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
