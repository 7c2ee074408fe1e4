python tools/time_flatten.py 4096
python tools/time_flatten.py 900
python tools/time_tiger.py
