# One gpurun call: parity of the front end and the search through the C ABI, serialized stage times of the bench-sized workload,
# and (optionally) a same-box A/B against another build of the library:  bash vosk-api_b200/tools/gpu_call.sh [path/to/other/libvosk.so]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tiny or small_model_all_stages or lattice_generation or max_active or full_size or pipelined" > gpurun_out/call_py.log 2>&1; tail -3 gpurun_out/call_py.log
for v in ${1:+alt} main; do
  if [ $v = alt ]; then export VOSK_B200_LIB=$1; else unset VOSK_B200_LIB; fi
  VB_SLOTS=1 timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "" 2 > gpurun_out/call_prof_$v.log 2>&1
  python - $v <<'PY'
import ast, sys
d = ast.literal_eval(open('gpurun_out/call_prof_%s.log' % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k: d[k] for k in ('ms_feat', 'ms_ivector', 'ms_nnet', 'ms_search', 'ms_prune')}, 'lane Mcycles', round(d['lane_cycles_sum'] / 1e6),
      {k[4:]: round(v / 1e6) for k, v in d.items() if k.startswith('cyc_') and v})
PY
done
