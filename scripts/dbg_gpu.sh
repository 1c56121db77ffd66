python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke ok|Error" | head -5
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40
