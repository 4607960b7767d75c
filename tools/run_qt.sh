for v in cl12 cl16 cl24; do python tools/quick_time.py variants/libdex_$v.so; done
