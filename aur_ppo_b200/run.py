#!/usr/bin/env python
"""`run.py` as README.md:10 and src/robot.sh:7-13 of the reference invoke it with the PPO flags
(-nm -nl -ne -d -do --t): those flags exist only in src/run_ppo.py, so this is that entry point."""
from .run_ppo import build_parser, main, params_from_args, strtobool  # noqa: F401

if __name__ == '__main__':
    main()
