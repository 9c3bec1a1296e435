"""CPU-side profile of bench.py (cProfile): where the host time of a step goes.
    python tools/prof_cpu.py [bench args...]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = ["bench.py"] + (sys.argv[1:] or ["--P", "2000", "--W", "64", "--H", "64", "--steps", "30", "--warmup", "3",
                                            "--no-cpu-baseline"])
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import bench  # noqa: E402

pr = cProfile.Profile()
pr.enable()
bench.main()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
