cd $GRAFT_REPO_ROOT
# compute-sanitizer record: the warp-per-robot kernels share shared memory and mbarriers between the warps of a
# CTA; the lane-per-robot kernels index a large global workspace.  Small batches: the tools serialise the kernels.
out=gpurun_out/sanitizer_r2.txt; : > $out
for tool in memcheck racecheck synccheck initcheck; do
  for what in warp lane rollout h30; do
    n=64; [ $tool = memcheck ] && n=256
    [ $tool != memcheck ] && [ $what = h30 ] && n=16
    echo "=== compute-sanitizer --tool $tool : san_driver.py $n $what" >> $out
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/san_driver.py $n $what 2>&1 | grep -v "^$" | tail -12 >> $out
  done
done
grep -c "ERROR SUMMARY: 0 errors" $out; grep "ERROR SUMMARY" $out | sort | uniq -c
