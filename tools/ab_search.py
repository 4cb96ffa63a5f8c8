"""Development helper: A/B timing of tuning builds of the search kernel.

    python tools/ab_search.py build            # here (nvcc): lib/libcellmapper_b200_<variant>.so for every variant
    python tools/ab_search.py run [shape ...]  # on the GPU box: tools/time_knn.py per variant, one JSON line each
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "base": (),
    "split": ("CM_SPLIT_EPI=1",),
}

if sys.argv[1] == "build":
    from concurrent.futures import ThreadPoolExecutor
    from cellmapper_b200 import build
    with ThreadPoolExecutor(max_workers=3) as pool:
        for name, path in zip(VARIANTS, pool.map(lambda kv: build.build(variant=kv[0], defines=kv[1]), VARIANTS.items())):
            print(name, path)
else:
    shapes = sys.argv[2:] or ["1500000x1500000x50"]
    for name in VARIANTS:
        lib = os.path.join(ROOT, "cellmapper_b200", "lib", f"libcellmapper_b200_{name}.so")
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, CM_LIBPATH=lib)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_knn.py"), *shapes], env=env, capture_output=True, text=True)
        for line in out.stdout.splitlines():
            if line.startswith("{"):
                print(json.dumps({"variant": name, **json.loads(line)}), flush=True)
        if out.returncode:
            print(json.dumps({"variant": name, "error": out.stderr[-400:]}), flush=True)
