for w in "6,2,7,64" "1,0,20,0" "12,2,7,64" "6,4,7,128" "6,2,5,64" "6,2,9,64" "3,1,6,32" "20,2,6,64" "6,1,7,16"; do echo "LPT $w"; PIXIE_LPT=$w python tools/time_tiger.py; done
