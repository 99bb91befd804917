timeout 600 python profiles/other_configs.py 2>&1 | grep -v Warning | tail -20
