python tools/quick_time.py
for v in mb10 mb12 nb128 nb32; do python tools/quick_time.py variants/libdex_$v.so; done
