# dev helper: run dev_c2 with an alternative build of the library (A/B experiments): CDBG_LIB=path
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcalm_b200 import api
if os.environ.get("CDBG_LIB"):
    api.LIB_PATH = os.environ["CDBG_LIB"]
sys.argv = [sys.argv[0]] + sys.argv[1:]
exec(open(os.path.join(os.path.dirname(__file__), "dev_c2.py")).read())
