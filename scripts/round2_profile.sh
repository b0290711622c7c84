#!/bin/bash
# round-2 evidence: ncu launch list of the bench command and two full captures (hot solve / Arnoldi kernels; assembly,
# factorisation and restart kernels).  The raw and per-kernel source pages are exported on the box (ncu is the same binary
# here and there) and only the CSVs travel back (gpurun_out is limited to 64 MiB).
mkdir -p gpurun_out
if [ "$1" != "skip-launches" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sharded > gpurun_out/bench_under_ncu_r2.log 2>&1
gzip -f gpurun_out/launches_r2.csv
fi
mkdir -p /tmp/ncu_r2
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:krylov_cgs2|slu_upper|slu_bwd_stage|slu_fwd_stage" -s 400 -c 8 -f -o /tmp/ncu_r2/hot python scripts/quick_bench.py 10001 > gpurun_out/ncu_full_hot.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:assemble_kernel|boundary_kernel|slu_build_rows|slu_merge|slu_top_factor|basis_gemm|bell_build" -c 22 -f -o /tmp/ncu_r2/setup python scripts/quick_bench.py 10001 > gpurun_out/ncu_full_setup.log 2>&1
for f in hot setup; do
  ncu -i /tmp/ncu_r2/$f.ncu-rep --page raw --csv > gpurun_out/full_r2_$f.raw.csv 2>/dev/null
done
for k in krylov_cgs2 slu_upper slu_bwd_stage slu_fwd_stage; do
  ncu -i /tmp/ncu_r2/hot.ncu-rep --page source --csv -k "regex:$k" -c 1 2>/dev/null | gzip > gpurun_out/src_r2_$k.csv.gz
done
for k in slu_merge assemble_kernel basis_gemm; do
  ncu -i /tmp/ncu_r2/setup.ncu-rep --page source --csv -k "regex:$k" -c 1 2>/dev/null | gzip > gpurun_out/src_r2_$k.csv.gz
done
ls -la gpurun_out | tail -16; du -sh gpurun_out
