import sys, time, torch
sys.path.insert(0, '/root/repo')
import tts_cube_b200 as cube
from oracle import wavernn_ref as R
dev = torch.device('cuda:0')
for (H, L, B, T_frames, up, upl, head) in [(512, 2, 20, 24, 100, 10, 'mol'), (512, 2, 1, 24, 100, 10, 'mol'), (512, 2, 20, 24, 100, 10, 'mulaw')]:
    S = {'mol': 30, 'mulaw': 256}[head]
    sd = R.random_state_dict(H, L, True, S, seed=5)
    v = cube.WaveRNNVocoder(L, H, up, upl, True, output=head).to(dev); v.load_state_dict(sd)
    mel = torch.rand(B, T_frames, 80, device=dev); xl = torch.rand(B, T_frames * up // upl, device=dev) * 1.6 - 0.8
    T = T_frames * up
    draws = v.make_draws(B, T, dev)
    for _ in range(2): x = v.inference(mel, xl, draws)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = v.inference(mel, xl, draws); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"wavernn H={H} L={L} head={head} B={B} T={T}: {dt*1e3:.1f} ms -> {T/dt:.0f} steps/s, {B*T/dt:.0f} samples/s ({B*T/dt/22050:.1f}x RT)")
    if B == 1:
        torch.set_num_threads(16)
        dr = {'u_mix': draws[:, :, :10].cpu(), 'u_x': draws[:, :, 10].cpu()}
        t0 = time.perf_counter(); R.wavernn_inference(sd, mel[:, :2].cpu(), xl[:, :20].cpu(), up, upl, head, {k: v_[:200] for k, v_ in dr.items()}); dt = time.perf_counter() - t0
        print(f"  CPU oracle (16 threads) B=1: {200/dt:.0f} samples/s")
