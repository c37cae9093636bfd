import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.argv = ["x"]
import tune
from vexcl_b200 import gen
cf = [("hell", 0, 0, 0), ("csr", 0, 2048, 512)]
tune.spmv_sweep(*tune.random_rows(4_000_000, 4_000_000, 12, 1), "random_avg12", cf)
tune.spmv_sweep(*tune.random_rows(1_000_000, 1_000_000, 60, 2), "random_avg60", cf)
tune.spmv_sweep(*gen.poisson_strip(3, 256), "poisson3d_256", cf[:1])
