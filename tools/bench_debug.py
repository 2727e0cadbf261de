"""Not a test: runs bench.py's main() with a watchdog that dumps every Python thread's stack if it does not finish."""
import faulthandler
import runpy
import sys

faulthandler.dump_traceback_later(int(sys.argv[1]), exit=True)
sys.argv = ["bench.py"] + sys.argv[2:]
runpy.run_path("bench.py", run_name="__main__")
