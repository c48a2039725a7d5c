# final evidence run of round 2 (one B200): GPU tests, smoke, bench line + reference arm (short), ncu launch list, traffic,
# --set full captures of the tensor-core, cluster and stationary-weights kernels, phase shares.  Outputs: gpurun_out/r3f_*.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3f_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r3f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r3f_bench.json
UIS_BENCH_REF_SECONDS=120 timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3f_ref.json 2> gpurun_out/r3f_ref.err; echo "ref rc=$?"; tail -c 300 gpurun_out/r3f_ref.json
python tools/prof_run.py 1 200 2 0 0 > gpurun_out/r3f_phases_stat_U1.txt 2>&1
python tools/prof_run.py 888 200 2 0 2 > gpurun_out/r3f_phases_tc_U888.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r3f_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:uis_beam_kernel -s 1 -c 1 --csv --log-file gpurun_out/r3f_traffic.csv python tools/prof_run.py 888 500 2 0 2 > gpurun_out/r3f_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uis_beam_kernel -s 1 -c 1 -o gpurun_out/r3f_beam_tc -f python tools/prof_run.py 888 60 2 0 2 > gpurun_out/r3f_prof_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uis_beam_kernel -s 1 -c 1 -o gpurun_out/r3f_beam_stat -f python tools/prof_run.py 1 60 2 0 0 > gpurun_out/r3f_prof_stat.log 2>&1
ls -la gpurun_out | grep r3f_
