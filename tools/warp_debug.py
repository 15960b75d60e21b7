import sys, numpy as np, torch
sys.path.insert(0, '.')
from hdpgpc_b200 import ops
from oracle import warp_oracle as W
z = np.load('tests/golden/warp_rec102_T90.npz')
x = z['x_basis'].reshape(-1); Y = z['data'][:, :, 0]
nb = z['noise_bounds']; n = float(np.clip(z['noise'], nb[0], nb[1]))
B = 40
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
scale = np.full(B, 1.0 / (B + 1e-12))
xw, yw, u, tr = ops.warp_fit_batched(cu(x), cu(Y[:B]), cu(Y[3][None]), n, 200.0, 1e-3, 8, 5e-2, 50, grad_scale=cu(scale), want_u=True, want_trace=True)
ref = z['fit_f_xw']
d = np.max(np.abs(xw[0].cpu().numpy() - ref), axis=1)
print('per-beat max diff', np.array2string(d, precision=2))
bad = int(np.argmax(d))
# oracle per-iteration per-beat loss
tab = W.ctrl_interp_table(8, 90)
uo = np.zeros((B, 8)); m = np.zeros_like(uo); v = np.zeros_like(uo)
for it in range(1, 51):
    terms, du, _, _ = W.loss_and_grad(uo, x, Y[:B], np.broadcast_to(Y[3], (B, 90)), tab, n, 200.0, 1e-3, scale)
    dl = abs(terms['loss'][bad] - float(tr[it - 1, 0, bad])) / abs(terms['loss'][bad])
    m = 0.9 * m + 0.1 * du; v = 0.999 * v + 0.001 * du * du
    uo = uo - (5e-2 / (1 - 0.9 ** it)) * m / (np.sqrt(v) / np.sqrt(1 - 0.999 ** it) + 1e-8)
    if it < 6 or dl > 1e-12:
        print(it, 'beat', bad, 'loss rel diff', dl, 'du', np.array2string(du[bad], precision=3))
        if dl > 1e-9: break
print('u diff', np.abs(u[0, bad].cpu().numpy() - uo[bad]))
