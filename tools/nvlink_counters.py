"""Sum of the NVLink data counters (KiB transmitted / received) per GPU from `nvidia-smi nvlink -gt d`; prints one JSON object."""
import json, re, subprocess, sys
out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True).stdout
res, gpu = {}, None
for line in out.splitlines():
    m = re.match(r"GPU (\d+):", line)
    if m:
        gpu = m.group(1); res[gpu] = {"tx_kib": 0, "rx_kib": 0}; continue
    m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
    if m and gpu is not None:
        res[gpu]["tx_kib" if m.group(1) == "Tx" else "rx_kib"] += int(m.group(2))
print(json.dumps(res))
