"""Full-depth golden fixture: the reference's UNMODIFIED files (through oracle/ref_shims) on the DEFAULT architecture
([8, 24, 24] coupling blocks, fc 1024; 61 nodes, 11.3 M parameters), B = 32.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_full.py        # needs /root/reference; writes tests/golden/full/full_depth.pt

Two initialisations:
  * `weight_init = 0.5` -- the well-conditioned full-depth config on which north_star's end-to-end tolerances are testable
    (reference fp32-vs-fp64 z error ~3e-5): z, log-det, per-sample losses, logits in fp64 (truth) and fp32 (the reference's
    own arithmetic), train- and eval-mode BatchNorm, plus d loss / d mu, d phi and the global gradient norm;
  * the default init (`weight_init = 1.0`), which is chaotic end to end (reference fp32-vs-fp64 z error ~1e-2): only z / jac in
    fp64 and fp32, so that a test can report err(engine) / err(reference fp32).
The 45 MB of parameters are NOT stored: the constructors are bit-identical under the same seeds
(tests/test_golden.py::test_construction_matches_reference); a per-tensor checksum pins that here.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'ref_shims'))
sys.path.insert(0, os.path.dirname(HERE))
import load_reference as lr            # noqa: E402

NO_DROP = {'dropout_conv': '[0., 0., 0.]', 'dropout_fc': '0.'}
B, BETA, SMOOTH = 32, 1.0, 0.02


def f32(t):
    return t.detach().to(torch.float32).clone()


def checksum(model):
    """(sum, sum of squares) per state_dict tensor in fp64"""
    return {k: (float(v.double().sum()), float((v.double() ** 2).sum())) for k, v in model.inn.state_dict().items()
            if 'tmp_var' not in k and v.is_floating_point()}


def main():
    ref = lr.load()
    out = {'B': B, 'beta': BETA}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, 32, 32, generator=g)
    lab = torch.randint(0, 10, (B,), generator=g)
    y = torch.zeros(B, 10).scatter_(1, lab.view(-1, 1), 1.) * (1 - SMOOTH) + SMOOTH / 10
    out['x'], out['y'] = x, y
    for tag, wi in (('wi05', '0.5'), ('default', '1.0')):
        ov = {'model': dict(NO_DROP, weight_init=wi)}
        args = lr.make_args(ov)
        torch.manual_seed(0)
        np.random.seed(0)
        m32 = ref.GenerativeClassifier(args)
        rec = {'overrides': ov, 'checksum': checksum(m32)}
        for dt, name in ((torch.float64, 'ref64'), (torch.float32, 'ref32')):
            for train in (True, False):
                if tag == 'default' and not train:
                    continue
                mm = copy.deepcopy(m32).to(dt)
                mm.train(train)
                xd, yd = x.to(dt), y.to(dt)
                res = mm(xd, yd, loss_mean=False)
                r = {k: f32(v) for k, v in res.items()}
                # z / jac: the forward above left them in the graph container
                mm2 = copy.deepcopy(m32).to(dt)
                mm2.train(train)
                with torch.no_grad():
                    r['z'] = f32(mm2.inn(xd))
                    r['jac'] = f32(mm2.inn.log_jacobian(run_forward=False))
                if tag == 'wi05' and train and dt == torch.float64:
                    bx, by = 2. / (1 + BETA), 2. * BETA / (1 + BETA)
                    (bx * res['L_x_tr'].mean() - by * res['L_y_tr'].mean()).backward()
                    r['g_mu'], r['g_phi'] = f32(mm.mu.grad), f32(mm.phi.grad)
                    r['grad_norm'] = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in mm.trainable_params if p.grad is not None)))
                    r['running'] = {k: f32(v) for k, v in mm.inn.state_dict().items() if 'running_' in k}
                rec['%s_%s' % (name, 'train' if train else 'eval')] = r
        out[tag] = rec
        e = (rec['ref32_train']['z'] - rec['ref64_train']['z']).abs().max() / rec['ref64_train']['z'].abs().max()
        print('%-8s reference fp32-vs-fp64 z rel err (train BN) %.2e' % (tag, e))
        if 'ref32_eval' in rec:
            e = (rec['ref32_eval']['z'] - rec['ref64_eval']['z']).abs().max() / rec['ref64_eval']['z'].abs().max()
            print('%-8s reference fp32-vs-fp64 z rel err (eval BN)  %.2e' % (tag, e))
    outdir = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'full')
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, 'full_depth.pt')
    torch.save(out, path)
    print(path, '%.2f MB' % (os.path.getsize(path) / 2 ** 20))


if __name__ == '__main__':
    main()
