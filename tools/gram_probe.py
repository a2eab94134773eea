"""Time the two products of the config-5 operator (A x and A^T w per shard) for the tuning knobs of gram.cu."""
import json
import os
import subprocess
import sys

code = r'''
import sys, json, torch
sys.path.insert(0, ".")
import arpack_ng_b200 as ab
G = ab.GramOperator.randsparse(20_000_000, 5_000_000, 16)
x = ab.hashed_start_vector(G.ncols)
y = torch.empty_like(x)
for _ in range(3): G(x, y)
torch.cuda.synchronize()
ab.profile(enable=True, reset=True)
for _ in range(10): G(x, y)
torch.cuda.synchronize()
p = ab.profile(enable=False)
print(json.dumps({k: (round(v["ms"] / v["launches"], 4), round(v["bytes"] / v["ms"] / 1e6)) for k, v in p.items() if k.startswith("gram")}))
'''
for env in ({"AB200_GRAM_LPR": "4", "AB200_GRAM_LD": "2"}, {"AB200_GRAM_LPR": "2", "AB200_GRAM_LD": "2"},
            {"AB200_GRAM_LPR": "1", "AB200_GRAM_LD": "2"}, {"AB200_GRAM_LPR": "2", "AB200_GRAM_LD": "2", "AB200_L2_WINDOW": "0"},
            {"AB200_GRAM_LPR": "4", "AB200_GRAM_LD": "2", "AB200_L2_WINDOW": "0"}, {"AB200_GRAM_LPR": "1", "AB200_GRAM_LD": "0"}):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    print(env, r.stdout.strip()[-300:], r.stderr.strip()[-200:], flush=True)
