timeout 200 python -m pytest tests -m gpu -q -x -s -k "gcn or synth3 or sample16 or sweep or config_shape or ragged" 2>&1 | grep -v Warn | grep "relative max\|passed\|failed\|Error" | head; 
timeout 100 python scripts/gcn_only.py 2048 2>&1 | tail -1
