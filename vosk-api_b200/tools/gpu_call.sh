# scratch driver of one gpurun call (edited per call): same-box A/B of builds of the library
for rep in 1 2; do
for v in A C main; do
  if [ $v != main ]; then export VOSK_B200_LIB=$PWD/vosk-api_b200/lib_alt/libvosk_$v.so; else unset VOSK_B200_LIB; fi
  VB_SLOTS=1 timeout 300 python vosk-api_b200/tools/profile_run.py 512 12 "" 2 > gpurun_out/e_prof1_$v$rep.log 2>&1
  python - $v$rep <<'PY'
import ast,sys
t=open('gpurun_out/e_prof1_%s.log'%sys.argv[1]).read().strip().splitlines()
d=ast.literal_eval(t[-1])
print(sys.argv[1], 'search', d['ms_search'], 'iv', d['ms_ivector'], 'hostlaunch', d['host_launch_ms'], 'max', d['lane_cycles_max'], 'sum', round(d['lane_cycles_sum']/1e6), {k[4:]:round(v/1e6) for k,v in d.items() if k.startswith('cyc_light') and v})
PY
done
done
