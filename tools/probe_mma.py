"""Development probe: time the tensor-core search kernel with parts of its pipeline disabled."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
import _probe_lib  # noqa: F401  (-DCM_DEV_PROBES build: the shipping library has no probe switches)
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cellmapper_b200 import _lib, device, synth

def run(n_q, n_r, d, flags, k=30, reps=3):
    centres = synth.mixture_centres(32, d)
    xr, _ = synth.mixture_embedding(n_r, centres, seed=1)
    xq, _ = synth.mixture_embedding(n_q, centres, seed=2)
    q = torch.from_numpy(xq).cuda(); r = torch.from_numpy(xr).cuda()
    lib = _lib.load(); lib.cm_profile_enable(1)
    buf = (ctypes.c_float * 4)()
    prof = torch.zeros(8 * 8192 + 8 + 8 * 8192, dtype=torch.int64, device='cuda')
    lib.cm_debug_probe_prof(prof.data_ptr())
    for f in flags:
        lib.cm_debug_probe_flags(f)
        out = []
        for i in range(reps + 1):
            prof.zero_()
            device.knn_search(q, r, k, dist_mode=_lib.DIST_SKLEARN_F32)
            lib.cm_profile_last_knn_ms(buf)
            if i: out.append(list(buf))
        ph = np.mean(out, 0)
        cd = prof.cpu().numpy()[8 * 8192:8 * 8192 + 8]
        tl = prof.cpu().numpy()[8 * 8192 + 8:].reshape(-1, 8); tl = tl[tl[:, 7] > 0]
        print('timeline (mean cycles since CTA start): setup %.0f  q_stored %.0f  first_acc %.0f  loop_done %.0f  final_compact %.0f  writeout %.0f  tiles %.1f  cta_end %.0f  (n=%d)' % (*tl.mean(0), len(tl)))
        pr = prof.cpu().numpy()[:8 * 8192].reshape(-1, 8); pr = pr[pr[:, 4] > 0]
        per = pr[:, :4].sum(0) / pr[:, 4].sum()
        ev = pr[:, 5:8].sum(0) / (4 * 2 * pr[:, 4].sum())  # per epilogue warp and tile (tiles counted for warp 1 = half)
        if cd[4] > 0: print('compaction per lane-call: iters %.2f  cycles minmax %.0f  search %.0f  rewrite %.0f  cnt %.1f  calls %d' % (cd[0]/cd[4], cd[1]/cd[4], cd[2]/cd[4], cd[3]/cd[4], cd[5]/cd[4], cd[4]))
        print(json.dumps(dict(n_q=n_q, n_r=n_r, d=d, flags=f, mma_ms=float(ph[1]), rerank_ms=float(ph[2]),
                              cyc_per_tile=dict(epi_slow_path_per_warp=per[0] / 8, of_which_compaction=per[1] / 8, issue=per[2], total=per[3]),
                              per_warp_tile=dict(triggered_halves=ev[0], drain_cycles=ev[1], compactions=ev[2]))), flush=True)
    lib.cm_debug_probe_flags(0)
    lib.cm_debug_probe_prof(None)

if __name__ == "__main__":
    shape = tuple(int(x) for x in sys.argv[1].split("x")) if len(sys.argv) > 1 else (100000, 100000, 50)
    flags = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3]
    run(*shape, flags)
