set -x
python tools/time_host_enqueue.py
