python tools/step_time.py 256 50
ORBX_CO=1 python tools/step_time.py 256 50
ORBX_CO=1 ORBX_FAST_PAD=8000 python tools/step_time.py 256 50
ORBX_CO=1 ORBX_FAST_PAD=16000 python tools/step_time.py 256 50
ORBX_CO=1 ORBX_FAST_PAD=16000 ORBX_BLUR_PAD=20000 python tools/step_time.py 256 50
ORBX_CO=1 ORBX_FAST_PAD=24000 ORBX_BLUR_PAD=20000 python tools/step_time.py 256 50
ORBX_CO=1 ORBX_FAST_PAD=24000 ORBX_BLUR_PAD=40000 python tools/step_time.py 256 50
ORBX_CO=1 ORBX_FAST_PAD=42000 ORBX_BLUR_PAD=20000 python tools/step_time.py 256 50
ORBX_FAST_PAD=8000 python tools/step_time.py 256 50
ORBX_FAST_PAD=16000 python tools/step_time.py 256 50
ORBX_FAST_PAD=24000 python tools/step_time.py 256 50
