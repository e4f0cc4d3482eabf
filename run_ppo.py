#!/usr/bin/env python
"""python run_ppo.py --gym_id CartPole-v1 ...  (the reference's src/run_ppo.py, B200 hot path)."""
from aur_ppo_b200.run_ppo import main

if __name__ == '__main__':
    main()
