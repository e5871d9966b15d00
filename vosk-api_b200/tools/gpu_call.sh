# scratch driver of one gpurun call (edited per call): full GPU suite and the bench line
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/j_py.log 2>&1; tail -3 gpurun_out/j_py.log
timeout 1500 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -c 300 gpurun_out/j_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02d_launch_list.csv python vosk-api_b200/tools/profile_run.py 512 4 > gpurun_out/j_ncu1.log 2>&1; tail -1 gpurun_out/j_ncu1.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_kernel" -s 4 -c 3 -f -o gpurun_out/r02d_prof_search python vosk-api_b200/tools/profile_run.py 512 4 > gpurun_out/j_ncu2.log 2>&1; tail -1 gpurun_out/j_ncu2.log | cut -c1-120
