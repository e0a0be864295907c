"""Quick GPU sanity run: synthetic configs -> CUDA path vs generator-expected PCM (and oracle).
`--small` leaves out the 1024-frame batch (for runs under compute-sanitizer)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import claxon_b200 as cb
from claxon_b200 import synth
from oracle import oracle as O

ctx = cb.Context(generic_only=('--generic' in sys.argv), warp_per_frame=('--warp' in sys.argv), lane_per_frame=('--warp' not in sys.argv))

def run(label, cfg):
    b = synth.generate(cfg)
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    t = time.time()
    out, res = ctx.decode_frames(b.data, descs, out_elems=out_elems)
    dt = time.time() - t
    bad_status = int((res["status"] != 0).sum())
    mism = 0
    for i in range(b.n_frames):
        o = int(descs[i]["out_offset"]); n = int(b.pcm_offsets[i+1] - b.pcm_offsets[i])
        if not np.array_equal(out[o:o+n], b.pcm[int(b.pcm_offsets[i]):int(b.pcm_offsets[i+1])]):
            mism += 1
            if mism <= 2:
                exp = b.pcm[int(b.pcm_offsets[i]):int(b.pcm_offsets[i+1])]
                w = np.nonzero(out[o:o+n] != exp)[0]
                print("   frame", i, "first diff at", w[:5], "got", out[o:o+n][w[:5]], "exp", exp[w[:5]], "status", res[i])
    cons_ok = bool((res["consumed"] == b.frame_lengths).all())
    print(f"{label}: frames={b.n_frames} samples={b.n_samples} bad_status={bad_status} mismatched_frames={mism} consumed_ok={cons_ok} e2e={dt*1e3:.1f}ms")
    if bad_status:
        print("   statuses:", np.unique(res["status"], return_counts=True))
    return mism == 0 and bad_status == 0 and cons_ok

ok = True
ok &= run("c2-64", synth.workload_config("c2", 64))
ok &= run("c2-indep", synth.workload_config("c2-indep", 64))
ok &= run("c3-64", synth.workload_config("c3", 64))
ok &= run("c4-110", synth.workload_config("c4", 110))
ok &= run("c5-4", synth.workload_config("c5", 4))
ok &= run("mixed", synth.SynthConfig(n_frames=200, block_size=1152, n_channels=2, bps=16, stereo_mode=-1, type_mask=15,
      lpc_min_order=1, lpc_max_order=32, qlp_precision=0, rice_mode=-2, rice_kmin=0, rice_kmax=14, max_porder=6,
      rice2=2, wasted_max=5, long_unary_per_mille=100))
ok &= run("ragged", synth.SynthConfig(n_frames=77, block_size=1000, tail_block_size=37, n_channels=3, bps=24, stereo_mode=0, type_mask=15,
      lpc_min_order=1, lpc_max_order=12, qlp_precision=0, rice_mode=-1, max_porder=3, wasted_max=3))
ok &= run("tiny", synth.SynthConfig(n_frames=50, block_size=16, tail_block_size=5, n_channels=2, bps=8, stereo_mode=-1, type_mask=15,
      lpc_min_order=1, lpc_max_order=16, qlp_precision=0, rice_mode=-1, max_porder=2, force_bs16=1))
if "--small" not in sys.argv:
    ok &= run("c2-1024", synth.workload_config("c2"))
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
