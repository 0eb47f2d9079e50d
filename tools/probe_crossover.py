"""Development probe: resident kernel against the streaming kernel on tissues between 512^2 and 2048^2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probe_mid import run

for n in (512, 640, 768, 896, 1024, 1200, 1536):
    for uni in (True, False):
        run(n, kernel=0, uniform=uni, steps=500)
        run(n, kernel=2, T=2, uniform=uni, steps=500)
        run(n, kernel=2, T=1, uniform=uni, steps=500)
