mkdir -p gpurun_out
MMU_TL_LIVE=1 python tools/step_timeline.py 10 > gpurun_out/r2h_timeline_live.json 2> gpurun_out/r2h_timeline_live.err; echo "exit $?"; tail -1 gpurun_out/r2h_timeline_live.err
MMU_TL_LIVE=1 MMU_EVAL_NOFOLD=1 MMU_BATTN_UNFUSED=1 python tools/step_timeline.py 10 > gpurun_out/r2h_timeline_live_old.json 2> gpurun_out/r2h_timeline_live_old.err; echo "exit $?"; tail -1 gpurun_out/r2h_timeline_live_old.err
python bench.py --live-tokens > gpurun_out/r2h_live.json 2> gpurun_out/r2h_live.err; echo "live exit $?"; head -c 300 gpurun_out/r2h_live.json; echo
