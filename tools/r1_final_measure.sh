set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r1f_tests.log
timeout 900 python bench.py > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
tail -c 600 gpurun_out/r1f_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r1f_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:uis_beam_kernel -s 1 -c 1 --csv --log-file gpurun_out/r1f_traffic.csv python tools/prof_run.py 296 500 2 2 > gpurun_out/r1f_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uis_beam_kernel -s 1 -c 1 -o gpurun_out/r1f_beam -f python tools/prof_run.py 296 60 2 2 > gpurun_out/r1f_prof.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_fit_launches.csv python tools/fit_prof.py 3 > gpurun_out/r1f_fit_prof.log 2>&1
ls -la gpurun_out | tail -12
