export PYTHONPATH=.
for cfg in "16 32 0" "16 16 0" "8 16 0" "8 32 0" "8 8 0" "16 8 0"; do set -- $cfg; echo "TZ=$1 CH=$2"; TOPAY_EDT_TZ=$1 TOPAY_EDT_CH=$2 KEEP_SQ=0 python scripts/field_probe.py 2>&1 | tail -2 | head -1; done
