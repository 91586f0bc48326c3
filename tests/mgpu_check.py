"""Multi-GPU parity (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mgpu_check.py

Every rank runs the hot path on its shard of a global batch with NCCL all-reduced statistics;
rank 0 also runs the whole batch alone.  Global normalisers => identical losses (to float32
reduction-order tolerance) and identical gradients / detections for the shard."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as tdist

from retinanet_tensorflow_b200 import dist, synthetic
from retinanet_tensorflow_b200.pipeline import DenseHeadPath


def run(wl, first, count, group, mode, logits=None, regs=None):
    boxes, classes, counts = synthetic.gt_batch(wl, count, first_image=first)
    path = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, count, wl.max_gt, cls_mode=mode,
                         process_group=group, capacity=8192, exchange=os.environ.get('RN_EXCHANGE', 'auto'))
    gb, gc, gn = (torch.from_numpy(x).cuda() for x in (boxes, classes, counts))
    if logits is None:
        single = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, count, wl.max_gt, process_group=False)
        single.assign(gb, gc, gn)
        logits, regs = synthetic.head_outputs(single.label, single.deltas, wl.num_classes, seed=100 + first,
                                              device='cuda')
    path.step(gb, gc, gn, logits, regs)
    torch.cuda.synchronize()
    # the pipelined form (detection tail + loss-sum exchange on the side stream, no host sync between steps) must
    # reproduce the serial step, also when the inputs change from step to step
    want = (path.loss_out.clone(), path.det_count.clone(), path.dlogits.clone())
    other = logits + 0.125
    for lg in (other, other, logits):
        path.step(gb, gc, gn, lg, regs, overlap_detect=True)
    path.join()
    torch.cuda.synchronize()
    if not (torch.equal(path.loss_out, want[0]) and torch.equal(path.det_count, want[1]) and
            torch.equal(path.dlogits, want[2])):
        raise SystemExit('MISMATCH overlapped step vs serial step (%s): %s vs %s' % (
            mode, path.loss_out.tolist(), want[0].tolist()))
    # the assignment launched one step ahead on its own stream (count exchange included) gives the same step
    for lg in (other, logits):
        path.step(gb, gc, gn, lg, regs, overlap_detect=True)
        path.prefetch_assign(gb, gc, gn)
    path.join()
    torch.cuda.synchronize()
    if not (torch.equal(path.loss_out, want[0]) and torch.equal(path.det_count, want[1]) and
            torch.equal(path.dlogits, want[2])):
        raise SystemExit('MISMATCH prefetched assignment vs serial step (%s): %s vs %s' % (
            mode, path.loss_out.tolist(), want[0].tolist()))
    path.check_status(ignore=12)
    # the whole multi-rank step as ONE CUDA graph (the exchange epochs live in device memory): replays reproduce it
    if path._peer is not None and path.exchange != 'nccl':
        gpath = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, count, wl.max_gt, cls_mode=mode,
                              process_group=group, capacity=8192, exchange='peer')
        graph = gpath.capture(gb, gc, gn, logits, regs)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        if not (torch.equal(gpath.loss_out, want[0]) and torch.equal(gpath.det_count, want[1]) and
                torch.equal(gpath.dlogits, want[2])):
            raise SystemExit('MISMATCH graph replay vs eager step (%s): %s vs %s' % (
                mode, gpath.loss_out.tolist(), want[0].tolist()))
        gpath.check_status(ignore=12)
    return path, logits, regs


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    tdist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    wl = synthetic.WORKLOADS['voc']
    global_batch = 2 * world + 1   # uneven split
    ok = True
    for mode in ('focal', 'bce_dice', 'balanced_bce'):
        first, count = dist.shard_images(global_batch, rank, world)
        path, logits, regs = run(wl, first, count, None, mode)
        # gather every shard's inputs on rank 0 and run the whole batch there
        shards = [None] * world
        tdist.all_gather_object(shards, (first, count, logits.cpu(), regs.cpu(), path.dlogits.cpu(),
                                         path.det_count.cpu(), path.loss_out.cpu()))
        if rank == 0:
            whole = DenseHeadPath((wl.image_h, wl.image_w), wl.num_classes, global_batch, wl.max_gt, cls_mode=mode,
                                  process_group=False, capacity=8192)
            boxes, classes, counts = synthetic.gt_batch(wl, global_batch)
            # merged level-major layout of the whole batch from the shards' merged buffers
            def merge(idx, width):
                parts = []
                for l in range(whole.pyr.num_levels):
                    for (f, c, *rest) in shards:
                        t = rest[idx - 2]
                        lv = DenseHeadPath.per_level_static(whole.pyr, c, t, width)[l]
                        parts.append(lv.reshape(-1, width))
                return torch.cat(parts, 0).cuda()
            lg, rg = merge(2, wl.num_classes), merge(3, 4)
            whole.step(*(torch.from_numpy(x).cuda() for x in (boxes, classes, counts)), lg, rg)
            torch.cuda.synchronize()
            want = whole.loss_out.cpu()
            for (f, c, _, _, dl, dc, lo) in shards:
                if not torch.allclose(lo, want, rtol=2e-6, atol=0):
                    ok = False
                    print('MISMATCH loss', mode, lo.tolist(), want.tolist())
            got_dl = merge(4, wl.num_classes)
            if not torch.allclose(got_dl, whole.dlogits, rtol=1e-6, atol=1e-12):
                ok = False
                print('MISMATCH dlogits', mode, float((got_dl - whole.dlogits).abs().max()))
            det = torch.cat([s[5] for s in shards])
            if not torch.equal(det, whole.det_count.cpu()):
                ok = False
                print('MISMATCH det_count', mode)
            print('mode %-13s world %d exchange %s: loss %s  (single-GPU %s)' % (
                mode, world, path.exchange if path._peer is None else 'peer', shards[0][6][:2].tolist(),
                want[:2].tolist()))
    flag = torch.tensor([1 if ok else 0], device='cuda')
    tdist.broadcast(flag, 0)
    tdist.barrier()
    tdist.destroy_process_group()
    if rank == 0:
        print('MGPU_CHECK', 'OK' if ok else 'FAILED')
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == '__main__':
    main()
