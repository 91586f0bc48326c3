cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tests/sanitize_check.py 2>&1 | grep -E "SANITIZE_RUN_DONE|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -12
done
timeout 900 python tests/fuzz_loss.py --cases 1500 --seed 21 2>&1 | tail -2
timeout 900 python tests/fuzz_parity.py --cases 400 --seed 21 2>&1 | tail -2
timeout 600 python tests/fuzz_select.py --cases 600 --seed 21 2>&1 | tail -2
