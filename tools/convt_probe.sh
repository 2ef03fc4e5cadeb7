# A/B timing of the epilogue layouts (MBS_EPI_VARIANT) on the layers the epilogue bounds
for v in 0 1; do for shp in 2,1,1024,1024,128,0,64 2,1,512,512,256,0,128 2,1,256,256,512,0,256 2,1,128,128,1024,0,512 0,1,2048,2048,64,0,64; do MBS_EPI_VARIANT=$v PROBE_ONLY=$shp PROBE_REPS=10 python tools/probe_conv.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print($v, d['mode'], d['H'],d['C0'],d['Cout'],d['ok'],round(d['ms'],4),round(d['tflops']))
"; done; done
