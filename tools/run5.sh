cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python tests/fuzz_loss.py --cases 600 --seed 5 2>&1 | tail -4
timeout 600 python tests/fuzz_select.py --cases 300 --seed 5 2>&1 | tail -3
