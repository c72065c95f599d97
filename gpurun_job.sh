python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/phase_times.py 64; python tools/phase_times.py 128; python tools/phase_times.py 96 1002 3
