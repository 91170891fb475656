import sys, argparse
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
a = argparse.Namespace(greedy_budget=1_000_000)
which = sys.argv[1]
if which == 'vecenv': bench.bench_vecenv(a)
elif which == 'ppo': bench.bench_ppo(a)
elif which == 'barcode': bench.bench_barcode(a)
elif which == 'pybase': bench.cpu_baseline_python(1 << 17, 36)
g = bench.bench_greedy(a)
print(which, g['seconds_wall'], g['seconds_device_max_group'], g['parity']['stored_paths_reproduced'])
