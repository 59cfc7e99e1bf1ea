timeout 120 python tools/newton_debug.py 203 7 2>&1 | tail -22
echo ---- poisoned workspace
timeout 120 python tools/newton_debug.py 203 4 empty 2>&1 | tail -22
