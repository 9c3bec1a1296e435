import cProfile, pstats, sys, os
sys.argv = ["bench.py", "--P", "2000", "--W", "64", "--H", "64", "--steps", "30", "--warmup", "3", "--no-cpu-baseline", "--streams", "4"]
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import bench
pr = cProfile.Profile()
pr.enable()
bench.main()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
