#!/bin/bash
# Attempt to pin the NF4 / nested-statistics oracle against real bitsandbytes 0.48.2 (the version the reference locks:
# /root/reference/uv.lock:307-308).  There is no network: the only source is the offline wheelhouse of the image.  If the
# install succeeds, tests/golden/make_bnb_golden.py dumps codes / absmax / nested pieces / dequantized weights for the
# golden inputs and tests/test_bnb_pin.py compares; if it fails, this log IS the record ("parity unpinned").
set -u
cd "$(dirname "$0")/.."
LOG=${1:-profiles/r02_bnb_pin_attempt.log}
{
  echo "# $(date -u +%FT%TZ) host=$(hostname) python=$(python -c 'import sys; print(sys.version.split()[0])')"
  echo "# wheelhouse candidates:"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'bitsandbytes|bnb' || echo "(none: no bitsandbytes wheel in /opt/wheelhouse)"
  echo "# importable already?"; python -c 'import bitsandbytes as b; print("bitsandbytes", b.__version__)' 2>&1 | tail -1
  echo "# pip install --no-index --find-links /opt/wheelhouse --target baseline/_ref bitsandbytes==0.48.2"
  python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref 'bitsandbytes==0.48.2' 2>&1 | tail -5
  echo "# pip download (index) -- expected to fail without network"
  timeout 20 python -m pip download --no-deps -d /tmp/bnb_dl 'bitsandbytes==0.48.2' 2>&1 | tail -2
  if PYTHONPATH=baseline/_ref python -c 'import bitsandbytes' 2>/dev/null; then
    echo "# bitsandbytes importable: dumping golden vectors"
    PYTHONPATH=baseline/_ref python tests/golden/make_bnb_golden.py
  else
    echo "# RESULT: bitsandbytes 0.48.2 cannot be installed here -> NF4 / nested parity stays UNPINNED (restatement of the published algorithm)"
  fi
} > "$LOG" 2>&1
cat "$LOG"
