"""Generate the committed golden vectors by executing the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Outputs tests/golden/*.npz.  Every fixture stores the seeds/sizes that regenerate its inputs
(rpnet_b200.synthetic + oracle.weights) and the reference outputs.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, weights                     # noqa: E402
from rpnet_b200.synthetic import make_episode, perturb_bn_stats  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings('ignore')
ref = ref_import.load()


def cfg_of(T, soft=False, radius=5):
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
                n_iter_refinement=T, soft_mask=soft, mask_refinement_correlation_radius=radius)


def build(T, soft=False, seed=0):
    torch.manual_seed(seed)                                 # test_rpnet.py:8-10
    net = ref.RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg_of(T, soft))
    return net


def run(net, ep):
    return net(ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'],
               query_labels=ep['query_labels'], appr_query_labels=ep['appr_query_labels'])


def save(name, **kw):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})
    print(name, '%.1f KB' % (os.path.getsize(path) / 1024))


# ---- init checksums -------------------------------------------------------------------------
net = build(1)
cs = weights.checksums(net.state_dict())
save('init_checksums', keys=np.array(list(cs.keys())), vals=np.array(list(cs.values())))
torch.manual_seed(0)
vgg = ref.Encoder(3)
cs = weights.checksums(vgg.state_dict())
save('init_checksums_vgg', keys=np.array(list(cs.keys())), vals=np.array(list(cs.values())))

# ---- cfg1: 1-shot 1-way, 2x128x128, T=1, eval (BASELINE.json configs[0]) -----------------------
feats = {}
net = build(1).eval()
perturb_bn_stats(net.state_dict(), seed=1)
net.encoder.register_forward_hook(lambda m, i, o: feats.setdefault('d4', []).append(o['d4']))
net.cre.register_forward_hook(lambda m, i, o: feats.setdefault('cre', []).append(o))
ep = make_episode(2, size=128, seed=0)
with torch.no_grad():
    out = run(net, ep)
assert torch.equal(out['output'], out['refinement'][0])
save('cfg1_eval', B=2, size=128, T=1, ep_seed=0, w_seed=0, bn_seed=1,
     output=out['output'], d4_supp=feats['d4'][0][:, ::8, ::2, ::2], d4_qry=feats['d4'][1][:, ::8, ::2, ::2],
     cre_supp=feats['cre'][0][:, ::4, ::2, ::2], cre_qry=feats['cre'][1][:, ::4, ::2, ::2])

# ---- T=3 hard mask and T=2 soft mask, 2x128x128 -------------------------------------------------
for name, T, soft in [('cfg1_T3', 3, False), ('cfg1_T2_soft', 2, True)]:
    net = build(T, soft).eval()
    perturb_bn_stats(net.state_dict(), seed=1)
    ep = make_episode(2, size=128, seed=10)
    with torch.no_grad():
        out = run(net, ep)
    kw = {'ref%d' % i: out['refinement'][i][:, :, ::2, ::2] for i in range(T)}
    kw.update({'mask%d' % i: np.packbits(out['refinement'][i].argmax(1).numpy().astype(np.uint8)) for i in range(T)})
    save(name, B=2, size=128, T=T, soft=soft, ep_seed=10, w_seed=0, bn_seed=1, **kw)

# ---- Correlation (net/rp_net.py:153-181) ----------------------------------------------------
g = torch.Generator().manual_seed(5)
kw = {}
for i, (b, c, h, w, r) in enumerate([(2, 16, 12, 10, 2), (1, 8, 16, 16, 5), (1, 4, 6, 7, 3)]):
    f1 = torch.randn(b, c, h, w, generator=g)
    f2 = torch.randn(b, c, h, w, generator=g)
    kw['f1_%d' % i], kw['f2_%d' % i], kw['r_%d' % i] = f1, f2, r
    kw['out_%d' % i] = ref.Correlation(f1, f2, r=r)
save('correlation', n=3, **kw)

# ---- getFeatures / getPrototype / calDist / alignLoss / dice_ce -----------------------------
net = build(1)
g = torch.Generator().manual_seed(7)
fts = torch.randn(1, 64, 16, 16, generator=g).relu()
mask = (torch.rand(1, 64, 64, generator=g) > 0.7).float()
empty = torch.zeros(1, 64, 64)
proto = net.getFeatures(fts, mask)
proto_empty = net.getFeatures(fts, empty)
qf = torch.randn(2, 64, 16, 16, generator=g).relu()
qf[0, :, 3, 4] = 0                                            # all-zero feature vector -> cosine 0
dist = net.calDist(qf, proto)
fg_l = [[torch.randn(1, 64, generator=g) for _ in range(3)] for _ in range(2)]
bg_l = [[torch.randn(1, 64, generator=g) for _ in range(3)] for _ in range(2)]
fgp, bgp = net.getPrototype(fg_l, bg_l)
logits = torch.randn(3, 2, 20, 24, generator=g) * 3
labels = (torch.rand(3, 20, 24, generator=g) > 0.6).long()
dce = ref.dice_ce(logits, labels)
logits5 = torch.randn(2, 5, 12, 12, generator=g)
labels5 = torch.randint(0, 5, (2, 12, 12), generator=g)
dce5 = ref.dice_ce(logits5, labels5)
# alignLoss: Wa=2, Sh=2
a_q = torch.randn(1, 64, 16, 16, generator=g).relu()
a_pred = torch.randn(1, 3, 16, 16, generator=g)
a_s = torch.randn(2, 2, 64, 16, 16, generator=g).relu()
a_f = (torch.rand(2, 2, 64, 64, generator=g) > 0.7).float()
a_b = 1 - a_f
a_b[0, 0, :5] = 0                                               # some ignore(255) pixels
al = net.alignLoss(a_q, a_pred, a_s, a_f, a_b)
save('proto_loss', fts=fts, mask=mask, proto=proto, proto_empty=proto_empty, qf=qf, dist=dist,
     fg_l=torch.stack([torch.stack(w) for w in fg_l]), bg_l=torch.stack([torch.stack(w) for w in bg_l]),
     fgp=torch.stack(fgp), bgp=bgp, logits=logits, labels=labels, dice_ce=dce, logits5=logits5, labels5=labels5,
     dice_ce5=dce5, a_q=a_q, a_pred=a_pred, a_s=a_s, a_f=a_f, a_b=a_b, align=al)

# ---- VGG encoder (net/vgg.py) standalone ------------------------------------------
torch.manual_seed(0)
vgg = ref.Encoder(3).eval()
x = make_episode(1, size=64, seed=3)['qry_imgs'][0].expand(-1, 3, -1, -1)
with torch.no_grad():
    y = vgg(x)
save('vgg', size=64, ep_seed=3, w_seed=0, out=y)

# ---- train step (SURVEY §3.5 reconstructed loss), 2x64x64, T=2 --------------------------------
net = build(2).train()
ep = make_episode(2, size=64, seed=20)
out = run(net, ep)
loss = 0
for i in sorted(out['refinement']):
    loss = loss + ref.dice_ce(out['refinement'][i], ep['query_labels'])
loss = loss + 1.0 * out['align_loss']
loss.backward()
names, gn, gs = [], [], []
for n, p in net.named_parameters():
    names.append(n)
    if p.grad is None:
        gn.append(-1.0); gs.append(np.zeros(4, np.float32))
    else:
        gn.append(p.grad.norm().item()); gs.append(p.grad.reshape(-1)[:4].numpy().copy() if p.grad.numel() >= 4
                                                   else np.resize(p.grad.reshape(-1).numpy(), 4))
bn = {k: v for k, v in net.state_dict().items() if 'running' in k or 'num_batches' in k}
save('train_step', B=2, size=64, T=2, ep_seed=20, w_seed=0, loss=loss, align=out['align_loss'],
     out0=out['refinement'][0], out1=out['refinement'][1], names=np.array(names), grad_norm=np.array(gn),
     grad_head=np.stack(gs), bn_keys=np.array(list(bn.keys())),
     bn_sums=np.array([v.double().sum().item() for v in bn.values()]))
print('done')
