timeout 900 python tools/crossover.py 150 257 400 600 900 1300 2>&1 | tail -12
