# round 2 / session 3 evidence run (one B200): ncu launch list of bench.py, DRAM/L2 traffic of the tensor-core beam kernel at
# bench size, --set full captures (with source) of the tensor-core kernel and of the cluster (latency-mode) kernel,
# per-phase cycle shares of both.  Outputs: gpurun_out/r3p_*; summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
python tools/prof_run.py 1 200 2 0 0 > gpurun_out/r3p_phases_cluster_U1.txt 2>&1
python tools/prof_run.py 888 200 2 0 2 > gpurun_out/r3p_phases_tc_U888.txt 2>&1
tail -3 gpurun_out/r3p_phases_cluster_U1.txt gpurun_out/r3p_phases_tc_U888.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r3p_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:uis_beam_kernel -s 1 -c 1 --csv --log-file gpurun_out/r3p_traffic.csv python tools/prof_run.py 888 500 2 0 2 > gpurun_out/r3p_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uis_beam_kernel -s 1 -c 1 -o gpurun_out/r3p_beam_tc -f python tools/prof_run.py 888 60 2 0 2 > gpurun_out/r3p_prof_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uis_beam_kernel -s 1 -c 1 -o gpurun_out/r3p_beam_cluster -f python tools/prof_run.py 1 60 2 0 0 > gpurun_out/r3p_prof_cluster.log 2>&1
ls -la gpurun_out | tail -12
