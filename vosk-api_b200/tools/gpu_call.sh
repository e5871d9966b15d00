# scratch driver of one gpurun call (edited per call): parity, then same-box A/B of two builds of the library
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tiny or small_model_all_stages or lattice_generation or max_active or full_size or pipelined or rule5 or partial or determin or silence or native or acoustic" > gpurun_out/k_py.log 2>&1; tail -3 gpurun_out/k_py.log
for rep in 1 2; do
for v in A main; do
  if [ $v != main ]; then export VOSK_B200_LIB=$PWD/vosk-api_b200/lib_alt/libvosk_$v.so; else unset VOSK_B200_LIB; fi
  VB_SLOTS=1 timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "" 2 > gpurun_out/k_prof1_$v$rep.log 2>&1
  timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "lattice=0" 3 > gpurun_out/k_prof_$v$rep.log 2>&1
  python - $v$rep <<'PY'
import ast,sys
t=open('gpurun_out/k_prof1_%s.log'%sys.argv[1]).read().strip().splitlines()
d=ast.literal_eval(t[-1])
u=open('gpurun_out/k_prof_%s.log'%sys.argv[1]).read().strip().splitlines()
print(sys.argv[1], 'search', d['ms_search'], 'max', d['lane_cycles_max'], 'sum', round(d['lane_cycles_sum']/1e6), {k[10:]:round(v/1e6) for k,v in d.items() if k.startswith('cyc_light') and v}, {k[10:]:round(v/1e6) for k,v in d.items() if k.startswith('cyc_heavy') and v}, '| best-path:', u[-2])
PY
done
done
