"""run pytest with an alternative build of the library: python run_with_lib.py <lib.so> <pytest args...>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icp-symm_b200"))
import icpsymm
icpsymm.LIB_PATH = sys.argv[1]
import pytest
sys.exit(pytest.main(sys.argv[2:]))
