"""Long randomized parity soak: VecEnv (all three ram kernels, image modes) vs the C oracle on the same Philox streams.
Not part of the test suite (minutes of CPU time); run on the GPU box:  python tests/manual/soak_parity.py [paths]
(paths: comma-separated subset of warp,cols,thread — default all; "cols" alone also skips the image configs)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import torch
import gym_simpletetris_b200 as st
from oracle.oracle import rollout
from _cases import CASES

INFO13 = [2, 0, 3, 4, 5, 7] + list(range(8, 15))
RAM_PATHS = sys.argv[1].split(",") if len(sys.argv) > 1 else ["warp", "cols", "thread"]
rs = np.random.RandomState(2026)
total = 0
t0 = time.time()
for name, kw in CASES.items():
    image = kw.get("obs_type", "ram") != "ram"
    n, T = (96, 600) if image else (1500, 2500)
    if image and len(sys.argv) > 1:
        continue
    for path in (["auto"] if image else RAM_PATHS):
        os.environ["ST_B200_RAM_PATH"] = path
        # biased action mix: more hard drops and rotations than uniform, to reach deeper stacks and more clears
        acts = rs.choice(7, size=(T, n), p=[0.16, 0.16, 0.22, 0.08, 0.14, 0.14, 0.10]).astype(np.uint8)
        env = st.VecEnv(n, device="cuda:0", seed=777, env_id_base=3, **kw)
        env.reset()
        a_dev = torch.from_numpy(acts).cuda()
        rew, don, inf = [], [], []
        for t in range(T):
            obs, r, d, info = env.step(a_dev[t])
            rew.append(r.clone()); don.append(d.clone()); inf.append(env.info_buf[:, INFO13].clone())
        got_r = torch.stack(rew).cpu().numpy(); got_d = torch.stack(don).cpu().numpy().astype(np.uint8)
        got_i = torch.stack(inf).cpu().numpy()
        okw = {k: v for k, v in kw.items() if k != "extend_dims"}
        want = rollout(n, acts, seed=777, env_id_base=3, want_info=True, **okw)
        assert np.array_equal(got_r, want["reward"]), (name, path, "reward")
        assert np.array_equal(got_d, want["done"]), (name, path, "done")
        assert np.array_equal(got_i, want["info"]), (name, path, "info")
        assert np.array_equal(obs.reshape(n, -1).cpu().numpy(), want["obs"]), (name, path, "obs")
        assert env.poll_errors() == 0
        total += n * T
        print(f"{name:28s} {path:6s} ok  episodes {int(got_d.sum()):6d}  lines {int(got_i[:, :, 3].max()):3d} max/episode  "
              f"clears>0 steps {int((np.diff(got_i[:, :, 3], axis=0) > 0).sum()):6d}", flush=True)
print(f"soak parity ok: {total / 1e6:.1f} M env-steps compared in {time.time() - t0:.0f} s")
