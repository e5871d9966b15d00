timeout 300 python vosk-api_b200/tools/gemm_selftest.py > gpurun_out/q_self.log 2>&1; tail -4 gpurun_out/q_self.log
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "tiny_model_all_stages or small_model_all_stages or large_arch" > gpurun_out/q_py.log 2>&1; tail -3 gpurun_out/q_py.log
timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "" 2 > gpurun_out/q_prof.log 2>&1; python - <<'PY'
import re,ast
t=open('gpurun_out/q_prof.log').read().strip().splitlines()
print(t[-2] if len(t)>1 else t)
d=ast.literal_eval(t[-1]); print({k:d[k] for k in ('ms_feat','ms_ivector','ms_nnet','ms_search')})
PY
VB_SLOTS=1 timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "" 2 > gpurun_out/q_prof1.log 2>&1; python - <<'PY'
import ast
t=open('gpurun_out/q_prof1.log').read().strip().splitlines()
print("serialized:", t[-2] if len(t)>1 else t)
d=ast.literal_eval(t[-1]); print({k:d[k] for k in ('ms_feat','ms_ivector','ms_nnet','ms_search')})
PY
