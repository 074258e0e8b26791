timeout 600 python -m pytest tests -q -m gpu -x -k "reprojection_error" 2>&1 | tail -12
