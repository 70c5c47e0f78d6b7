"""Golden vectors produced by the REFERENCE's own code, run in the build container (it cannot run on the GPU box:
/root/reference does not travel; casadi / l4casadi / acados / adam are not installable offline).

What of the hot path can be executed from /root/reference without those packages is the viability network itself:
``safe_mpc.safe_set.NeuralNetwork`` (safe_set.py:26-43) is plain torch.  The module imports casadi and l4casadi at its
top, so two empty stub modules are put in ``sys.modules`` for the import; the class body is the reference's, unmodified.
The weights are the ones the engine loads (``problem.load_network``: the reference's file format, safe_set.py:76-85), the
activation is the one reference parser.py:96-103 maps 'gelu' to (``GELU(approximate='tanh')``).

Stored: inputs x, psi(x) as safe_set.py:82-87 defines it (restated here in torch so that autograd goes through it), the network
output in fp32 (what the reference's L4CasADi call computes) and in fp64 (the same class, ``.double()``), the constraint
c(x) = NN(psi) (100 - alpha)/100 - |v| (safe_set.py:100-104) and dc/dx by autograd through the reference network.

    python tests/golden/make_ref_golden.py        ->  tests/golden/ref_network.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference/src'


def reference_network_class():
    for name in ('casadi', 'l4casadi'):                     # imported at the top of safe_set.py, not used by NeuralNetwork
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_safe_set', os.path.join(REF, 'safe_mpc', 'safe_set.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.NeuralNetwork


def main():
    from tests.common import params_model, random_states, start_states
    from safe_mpc_b200.problem import load_network
    params, md = params_model(alpha=10.0)
    ws, bs, mean, std = load_network(params)
    Net = reference_network_class()
    net = Net(2 * params.nq, int(params.net_size[1]), 1, torch.nn.GELU(approximate='tanh'))     # parser.py:100, env_model / safe_set.py:76-81
    sd = {}
    for i, li in enumerate((0, 2, 4, 6)):
        sd[f'linear_stack.{li}.weight'] = torch.tensor(ws[i]); sd[f'linear_stack.{li}.bias'] = torch.tensor(bs[i])
    net.load_state_dict(sd)
    net.eval()
    x = np.vstack([random_states(md, 96, seed=5, vel_scale=0.5), start_states(32, seed=6, vel=0.4)])
    nq, eps, alpha = params.nq, float(params.eps), float(params.alpha)

    def psi(xt, dtype):                                      # safe_set.py:82-87
        q, v = xt[:, :nq], xt[:, nq:].clone()
        v[:, 0] = v[:, 0] + eps
        nrm = torch.linalg.norm(v, dim=1, keepdim=True)
        return torch.cat([(q - torch.tensor(mean, dtype=dtype)) / torch.tensor(std, dtype=dtype), v / nrm], dim=1), nrm[:, 0]

    with torch.no_grad():
        p32, _ = psi(torch.tensor(x, dtype=torch.float32), torch.float32)
        y32 = net(p32)[:, 0].numpy()
    net64 = net.double()
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    p64, nrm = psi(xt, torch.float64)
    y64 = net64(p64)[:, 0]
    c = y64 * (100.0 - alpha) / 100.0 - nrm                  # safe_set.py:100-104
    grad = torch.autograd.grad(c.sum(), xt)[0]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_network.npz')
    np.savez_compressed(out, x=x, psi=p64.detach().numpy(), y32=y32, y64=y64.detach().numpy(), c=c.detach().numpy(), grad=grad.numpy(),
                        alpha=alpha, eps=eps)
    print('wrote', out, 'max |y32 - y64| =', float(np.abs(y32 - y64.detach().numpy()).max()))


if __name__ == '__main__':
    main()
