# round 2 measurement run (one B200): bench line, reference arm smoke, ncu launch list / traffic / full capture of the
# tensor-core beam kernel.  Outputs under gpurun_out/r2m_*; the summaries are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2m_bench.json; tail -5 gpurun_out/r2m_bench.err
UIS_BENCH_REF_SECONDS=90 timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_ref.json 2> gpurun_out/r2m_ref.err; echo "ref rc=$?"
tail -c 1200 gpurun_out/r2m_ref.json; tail -3 gpurun_out/r2m_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2m_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:uis_beam_kernel -s 1 -c 1 --csv --log-file gpurun_out/r2m_traffic.csv python tools/prof_run.py 888 500 2 0 2 > gpurun_out/r2m_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uis_beam_kernel -s 1 -c 1 -o gpurun_out/r2m_beam_tc -f python tools/prof_run.py 888 60 2 0 2 > gpurun_out/r2m_prof.log 2>&1
ls -la gpurun_out | tail -12
