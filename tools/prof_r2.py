"""Small profiling targets for ncu (round 2): `cov [batch] [steps]` runs key-collection steps
(fast path to layer 8 + second-moment col-GEMM); `insert [niter]` runs the fused insert loop."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import bench
    from rewriting_b200 import fastpath
    from rewriting_b200.utils import runningstats, zdataset
    mode = sys.argv[1]
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    model = bench.build_model(dev)
    if mode == 'cov':
        batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
        steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
        z = zdataset.standard_z_sample(batch * steps, 512, seed=1).to(dev)
        r2m = runningstats.RunningSecondMoment()
        with torch.no_grad():
            for i in range(steps):
                planes = fastpath.forward(model, z[i * batch:(i + 1) * batch], upto_key_layer=8)
                r2m.add_planes(planes.hi, planes.lo, planes.B * planes.H * planes.W)
        torch.cuda.synchronize()
        print('cov ok', r2m.count)
    elif mode == 'insert':
        niter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
        z = zdataset.standard_z_sample(32, 512, seed=1).to(dev)
        print(bench.bench_insert(model, z, dev, niter=niter))


if __name__ == '__main__':
    main()
