# scratch driver of one gpurun call (edited per call)
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/b_py.log 2>&1; tail -4 gpurun_out/b_py.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/b_launches.csv python vosk-api_b200/tools/profile_run.py 512 4 > gpurun_out/b_ncu1.log 2>&1; tail -1 gpurun_out/b_ncu1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mfcc|ivector" -s 8 -c 8 -f -o gpurun_out/r02b_prof_fe python vosk-api_b200/tools/profile_run.py 512 4 > gpurun_out/b_ncu2.log 2>&1; tail -1 gpurun_out/b_ncu2.log | cut -c1-200
