python tools/time_tiger.py
for v in 4 8 12; do PIXIE_CUDA_LIB=build/ab/lm$v.so python tools/time_tiger.py; done
for v in 4 8 12; do PIXIE_CUDA_LIB=build/ab/lm$v.so python tools/time_icons.py; done
python tools/time_icons.py
