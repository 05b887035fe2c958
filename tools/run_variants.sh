timeout 600 python -m pytest tests/test_frame.py tests/test_soak.py -m gpu -x -q 2>&1 | tail -3
echo "== 1x64 recorded"; timeout 200 python tools/time_frame.py 1 64 3 intra,all record
