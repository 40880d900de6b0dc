#!/bin/bash
# ncu evidence: launch list (time + DRAM bytes) of the bench command, full captures of the top kernels,
timeout 600 python -m pytest tests/test_gpu_graph.py -q -x -k "balanc" > gpurun_out/n_pytest.log 2>&1; tail -3 gpurun_out/n_pytest.log
# DRAM/time of the CPCA and hub-paring kernels
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-check"
$BENCH > gpurun_out/n_plain_bench.log 2>&1 && \
ncu --metrics $M --clock-control none -c 12000 --csv --log-file gpurun_out/n_launches_bench.csv $BENCH > gpurun_out/n_ncu_bench.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/n_ncu_bench.log
PROF="python tools/profile_stages.py --config 2 --reps 0"
$PROF > gpurun_out/n_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:knn_tc2 -c 2 -o gpurun_out/n_knn_tc2 $PROF > gpurun_out/n_ncu_knn.log 2>&1
echo "knn full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_bf16x3 -c 1 -o gpurun_out/n_gemm_gram $PROF > gpurun_out/n_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none -k "regex:eig_rr_kernel|eig_orth_kernel" -c 2 -o gpurun_out/n_eig $PROF > gpurun_out/n_ncu_eig.log 2>&1
echo "eig full rc=$?"
CP="python tools/profile_stages.py --config 3 --n-samples 8 --reps 0"
$CP > gpurun_out/n_plain_cpca.log 2>&1 && \
ncu --set full --clock-control none -k "regex:skinny_gemm_kernel|cpca_step_kernel" -s 40 -c 2 -o gpurun_out/n_cpca $CP > gpurun_out/n_ncu_cpca.log 2>&1
echo "cpca rc=$?"
HB="python tools/profile_stages.py --config 5 --n-samples 2 --reps 0"
$HB > gpurun_out/n_plain_hub.log 2>&1 && \
ncu --metrics $M,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k "regex:pare_kernel|knn_tcl" -c 12 --csv --log-file gpurun_out/n_launches_hub.csv $HB > gpurun_out/n_ncu_hub.log 2>&1
echo "hub rc=$?"
ls -la gpurun_out/n_* | head -30
