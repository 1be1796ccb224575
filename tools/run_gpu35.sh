set -x
timeout 100 python tools/time_qr.py 16384 2>&1 | head -1
GLA_PANEL_MINROWS=128 timeout 100 python tools/time_qr.py 16384 2>&1 | head -1
GLA_PANEL_MINROWS=256 timeout 100 python tools/time_qr.py 16384 2>&1 | head -1
GLA_PANEL_MAXCTAS=32 GLA_PANEL_MINROWS=128 timeout 100 python tools/time_qr.py 16384 2>&1 | head -1
GLA_PANEL_MAXCTAS=96 timeout 100 python tools/time_qr.py 16384 2>&1 | head -1
