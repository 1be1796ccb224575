GLA_BATCHED_MODE=2 python tools/time_batched.py
GLA_BATCHED_MODE=3 python tools/time_batched.py
