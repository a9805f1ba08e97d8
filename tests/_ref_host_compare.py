"""Helper of tests/test_host_vs_reference.py — runs in a SUBPROCESS (the reference shim
monkey-patches torch globally): imports the live reference next to this package and compares
the device-independent host helpers on seeded random inputs.  Prints one JSON object
{check name: max abs difference or 0/1 flag}."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference                 # noqa: E402

ref = load_reference()
from rewriting_b200.rewrite import ganrewrite as mine_gw    # noqa: E402
from rewriting_b200.utils import nethook as my_nethook, renormalize as my_renorm  # noqa: E402
from rewriting_b200.utils import zdataset as my_z           # noqa: E402
from rewriting_b200.utils.sampler import FixedSubsetSampler as MySampler  # noqa: E402

out = {}
rng = np.random.RandomState(0)
g = torch.Generator().manual_seed(0)


def diff(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    if a.shape != b.shape:
        return float('inf')
    return float((a - b).abs().max()) if a.numel() else 0.0


# ---- zdataset -------------------------------------------------------------------------
out['z_sample'] = max(diff(my_z.standard_z_sample(n, d, seed=s), ref.zdataset.standard_z_sample(n, d, seed=s))
                      for n, d, s in [(5, 512, 1), (37, 64, 10), (1, 512, 20)])
out['y_sample'] = diff(my_z.standard_y_sample(50, 10, seed=3), ref.zdataset.standard_y_sample(50, 10, seed=3))

# ---- renormalize ------------------------------------------------------------------------
img = torch.rand(3, 40, 52, generator=g) * 2 - 1
worst = 0.0
for src in ('zc', 'pt', 'imagenet', 'byte'):
    for tgt in ('zc', 'pt', 'imagenet', 'byte'):
        x = img if src == 'zc' else ref.renormalize.as_tensor(img, 'zc', src)
        x = x.float()
        worst = max(worst, diff(my_renorm.as_tensor(x, src, tgt).float(), ref.renormalize.as_tensor(x, src, tgt).float()))
out['renorm_as_tensor'] = worst
out['renorm_as_image'] = diff(np.asarray(my_renorm.as_image(img)), np.asarray(ref.renormalize.as_image(img)))
url = ref.renormalize.as_url(img)
out['renorm_url_roundtrip'] = float(my_renorm.as_url(img) == url)
out['renorm_from_url'] = max(diff(my_renorm.from_url(url, target=t, size=sz), ref.renormalize.from_url(url, target=t, size=sz))
                             for t, sz in [('zc', None), ('pt', (32, 32)), ('byte', (16, 24))])

# ---- rewriter geometry helpers ------------------------------------------------------------
worst = {k: 0.0 for k in ('bbox', 'center', 'paste', 'crop')}
for trial in range(25):
    h, w = int(rng.randint(6, 40)), int(rng.randint(6, 40))
    mask = torch.zeros(h, w)
    t, l = int(rng.randint(0, h - 2)), int(rng.randint(0, w - 2))
    b, r = int(rng.randint(t + 1, h + 1)), int(rng.randint(l + 1, w + 1))
    mask[t:b, l:r] = torch.rand(b - t, r - l, generator=g) + 0.01
    worst['bbox'] = max(worst['bbox'], diff(mine_gw.positive_bounding_box(mask), ref.ganrewrite.positive_bounding_box(mask)))
    worst['center'] = max(worst['center'], diff(mine_gw.centered_location(mask), ref.ganrewrite.centered_location(mask)))
    src = torch.randn(1, 4, h, w, generator=g)
    ch, cw = int(rng.randint(1, h + 1)), int(rng.randint(1, w + 1))
    clip = torch.randn(1, 4, ch, cw, generator=g)
    area = torch.rand(ch, cw, generator=g)
    center = (int(rng.randint(0, h)), int(rng.randint(0, w)))
    for ar in (None, area):
        a1, b1 = mine_gw.paste_clip_at_center(src, clip, center, ar)
        a2, b2 = ref.ganrewrite.paste_clip_at_center(src, clip, center, ar)
        worst['paste'] = max(worst['paste'], diff(a1, a2), diff(b1, b2))
    tgt = torch.randn(1, 4, 2 * h, 2 * w, generator=g)
    bounds = (t, l, b, r)
    c1 = mine_gw.crop_clip_to_bounds(src, tgt, bounds)
    c2 = ref.ganrewrite.crop_clip_to_bounds(src, tgt, bounds)
    worst['crop'] = max(worst['crop'], max(diff(x, y) for x, y in zip(c1, c2)))
out.update({'geom_' + k: v for k, v in worst.items()})

# ---- zca_from_cov -------------------------------------------------------------------------
a = torch.randn(200, 24, generator=g)
cov = a.t() @ a / 200
out['zca_from_cov'] = diff(mine_gw.zca_from_cov(cov), ref.ganrewrite.zca_from_cov(cov))

# ---- nethook.subsequence on a nested Sequential ---------------------------------------------
from collections import OrderedDict  # noqa: E402


def toy():
    torch.manual_seed(3)
    return torch.nn.Sequential(OrderedDict([
        ('a', torch.nn.Linear(6, 6)),
        ('b', torch.nn.Sequential(OrderedDict([('b1', torch.nn.Linear(6, 6)), ('b2', torch.nn.Tanh()),
                                               ('b3', torch.nn.Linear(6, 6))]))),
        ('c', torch.nn.ReLU()), ('d', torch.nn.Linear(6, 3))]))


x = torch.randn(5, 6, generator=g)
worst, same_names, shared = 0.0, 1.0, 1.0
for kw in [dict(first_layer='b.b2', last_layer='c'), dict(after_layer='a', upto_layer='b.b3'),
           dict(first_layer='b', last_layer='b'), dict(upto_layer='b.b2'), dict(after_layer='b.b1')]:
    m1, m2 = toy(), toy()
    s1 = my_nethook.subsequence(m1, share_weights=True, **kw)
    s2 = ref.nethook.subsequence(m2, share_weights=True, **kw)
    inp = x
    worst = max(worst, diff(s1(inp).detach(), s2(inp).detach()))
    same_names = min(same_names, float([n for n, _ in s1.named_modules()] == [n for n, _ in s2.named_modules()]))
    ids = {id(p) for p in m1.parameters()}
    shared = min(shared, float(all(id(p) in ids for p in s1.parameters())))
out['subsequence_output'] = worst
out['subsequence_names_equal'] = same_names
out['subsequence_shares_weights'] = shared

# ---- InstrumentedModel retain / edit ----------------------------------------------------------
m1, m2 = toy(), toy()
i1, i2 = my_nethook.InstrumentedModel(m1), ref.nethook.InstrumentedModel(m2)
i1.retain_layers(['b.b1', ('d', 'out')])
i2.retain_layers(['b.b1', ('d', 'out')])
y1, y2 = i1(x), i2(x)
out['imodel_retained'] = max(diff(i1.retained_layer('b.b1').detach(), i2.retained_layer('b.b1').detach()),
                             diff(i1.retained_layer('out').detach(), i2.retained_layer('out').detach()),
                             diff(y1.detach(), y2.detach()))
rep = torch.randn(5, 6, generator=g)
for im in (i1, i2):
    im.edit_layer('b.b1', ablation=0.5, replacement=rep)
out['imodel_edit'] = diff(i1(x).detach(), i2(x).detach())
for im in (i1, i2):
    im.remove_edits()
out['imodel_edits_removed'] = diff(i1(x).detach(), i2(x).detach())
i1.close()          # (the reference's own close() trips an assert when a layer was retained under
                    #  an alias, so only this package's is exercised)
out['imodel_closed_restores'] = diff(m1(x).detach(), toy()(x).detach())

# ---- sampler ------------------------------------------------------------------------------------
from utils.sampler import FixedSubsetSampler as RefSampler   # noqa: E402
out['sampler'] = float(list(MySampler([3, 1, 4, 1, 5])) == list(RefSampler([3, 1, 4, 1, 5])) and
                       len(MySampler(list(range(7)))) == len(RefSampler(list(range(7)))))

print('RESULT ' + json.dumps(out))
