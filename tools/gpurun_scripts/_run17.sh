cd $GRAFT_REPO_ROOT
compute-sanitizer --tool memcheck --print-limit 20 python tools/diag_small.py 192 100 2>&1 | grep -vE "^=+$" | tail -40
compute-sanitizer --tool racecheck --print-limit 10 python tools/diag_small.py 160 30 2>&1 | grep -vE "^=+$" | tail -25
compute-sanitizer --tool initcheck --print-limit 10 python tools/diag_small.py 160 30 2>&1 | grep -vE "^=+$" | tail -25
