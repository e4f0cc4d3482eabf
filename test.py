"""`python test.py [model.pt]`: the reference's src/test.py entry point on the device envs (aur_ppo_b200/test.py)."""
import runpy
import sys

if __name__ == "__main__":
    sys.argv[0] = "aur_ppo_b200.test"
    runpy.run_module("aur_ppo_b200.test", run_name="__main__")
