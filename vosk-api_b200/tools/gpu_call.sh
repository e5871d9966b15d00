# scratch driver of one gpurun call (edited per call)
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tiny_model_all_stages or small_model_all_stages or large_arch or odd or kaldi or rule5 or silence" > gpurun_out/h_py.log 2>&1; tail -3 gpurun_out/h_py.log
VB_SLOTS=1 timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "" 2 > gpurun_out/h_prof1.log 2>&1; tail -1 gpurun_out/h_prof1.log | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ivector|mfcc" -c 40 --csv --log-file gpurun_out/h_launches.csv python vosk-api_b200/tools/profile_run.py 512 4 > gpurun_out/h_ncu1.log 2>&1; python vosk-api_b200/tools/launch_shares.py gpurun_out/h_launches.csv
