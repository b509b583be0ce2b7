set -x
mkdir -p gpurun_out
for m in 1 2; do
MMDFN_WGRAD_MODE=$m timeout 200 python tools/step_profile.py --graph > gpurun_out/r2t_prof_m$m.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2t_prof_m$m.log; cp gpurun_out/step_timeline.txt gpurun_out/r2t_timeline_m$m.txt
done
rm -f gpurun_out/step_trace.json
