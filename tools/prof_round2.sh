set -x
export ONLY="sa1.1 l3" REPS=2 NOGRAPH=1
python tools/microbench_gemm.py > gpurun_out/s9_plain_gemm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_rows -c 12 -o gpurun_out/r2_gemm_sa1 python tools/microbench_gemm.py > gpurun_out/s9_ncu_gemm1.log 2>&1
export ONLY="sa3.1 l3"
python tools/microbench_gemm.py > gpurun_out/s9_plain_gemm2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_rows -c 12 -o gpurun_out/r2_gemm_sa3 python tools/microbench_gemm.py > gpurun_out/s9_ncu_gemm2.log 2>&1
unset ONLY REPS NOGRAPH
python tools/bench_configs.py --only c1 --iters 1 > gpurun_out/s9_plain_c1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"sa_fused|gemm_rows" -s 60 -c 14 -o gpurun_out/r2_infer_c1 python tools/bench_configs.py --only c1 --iters 1 > gpurun_out/s9_ncu_c1.log 2>&1
python tools/ncu_step.py > gpurun_out/s9_plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_final_step.csv python tools/ncu_step.py > gpurun_out/s9_ncu_step.log 2>&1
ls -la gpurun_out/*.ncu-rep
