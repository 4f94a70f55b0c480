import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import fullsize_util as U
from oracle import nerv_oracle as O
dev = torch.device('cuda:0')
g = torch.load('/root/repo/tests/golden/small_erb.pt', map_location='cpu', weights_only=False)
def run(tag):
    pe, gen = U.build(g['cfg'], 'ERB', dev)
    embed = pe(g['pos'])
    img = gen(embed)[0]
    torch.cuda.synchronize()
    print(tag, 'img rel', U.rel_l2(img, g['img']), 'finite', bool(torch.isfinite(img).all()))
    ex = gen.executor(2, True)
    for i, (K_ref, b_ref) in enumerate(g['folded']):
        Kt = ex.K[i].view(K_ref.shape[0], 9, K_ref.shape[1]).permute(0, 2, 1).reshape(K_ref.shape)
        print('   block', i, 'K rel', U.rel_l2(Kt, K_ref), 'b rel', U.rel_l2(ex.bias[i], b_ref))
    feats = O.generator_forward({k: v for k, v in g['init_state'].items()}, O.pos_encoding(g['pos'], 1.25, 40), U.ocfg(g['cfg']), return_features=True)[1]
    for l in range(len(ex.x)):
        xr = feats[l].permute(0, 2, 3, 1)
        xo = ex.x[l].float()[..., :xr.shape[-1]]
        print('   x', l, 'rel', U.rel_l2(xo, xr))
run('fresh')
which = sys.argv[1] if len(sys.argv) > 1 else 'S720'
res = U.one_step_parity(which, dev)
print('big parity', res['img_rel_l2'])
run('after big')
run('again')
