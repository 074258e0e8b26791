timeout 600 python -m pytest tests -q -m gpu -x -s -k "many_clusters or venice" 2>&1 | tail -5
