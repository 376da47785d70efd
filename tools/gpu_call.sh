mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke exit $?"; grep -E "smoke\[|Error" gpurun_out/r2n_smoke.log
