python tools/time_tiger.py
for v in lb5 lb6; do PIXIE_CUDA_LIB=build/ab/$v.so python tools/time_tiger.py; done
python tools/time_icons.py | tail -1
for v in lb5 lb6; do PIXIE_CUDA_LIB=build/ab/$v.so python tools/time_icons.py | tail -1; done
