import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tools import probe_small as P
name, kernel, T, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
work = {"fk128": bench.make_fk128, "fk512": bench.make_fk512}[name]()
P.run(name, work, kernel, T, n=n)
