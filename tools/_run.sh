timeout 600 python -m pytest tests -q -m gpu -x -s -k "long_tracks" 2>&1 | tail -12
