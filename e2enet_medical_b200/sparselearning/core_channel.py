"""Drop-in replacement for e2enet/training/network_training/sparselearning/core_channel.py.

Same public surface as the reference (`add_sparse_args`, `CosineDecay`, `LinearDecay`,
`Masking(optimizer, death_rate, ..., args)` with `add_module / init / step / apply_mask /
truncate_weights / kernel_death / kernel_growth / cal_nonzero_counts / fired_masks_update /
print_nonzero_counts / death_decay_update` and the same public state), so
`simple_main.py:163-168` and `run_iteration` (`mask.step()`) run unchanged.

What differs is where the work happens: apply_mask is ONE multi-tensor CUDA kernel over all
35 masked tensors (reference: 70 allocating torch kernels), prune = kernel-L1 + exact radix
select + threshold kill per tensor without host syncs, regrow = ordered dead-list compaction
on the device + a scatter of the indices that Python's `random.sample` draws on the host
(same call sequence as the reference, so the RNG stream -- and therefore the index sets --
are bit-identical), and all counters come back in 3 device->host reads per update instead
of ~250 `.item()` calls.  Masks are updated IN PLACE so the conv kernels' weight packing
(which multiplies by the mask on load) always sees the current topology.
"""
from __future__ import print_function

import copy
import ctypes as C
import math
import random

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from .. import _lib


def str2bool(str):
    return True if str.lower() == 'true' else False


def add_sparse_args(parser):
    # identical flags / defaults, including the type=bool quirk of --adv / --fix (any non-empty string is True)
    parser.add_argument('--sparse', type=str2bool, default=True, help='Enable sparse mode. Default: True.')
    parser.add_argument('--adv', type=bool, default=False, help='adv sparse mode. Default: True.')
    parser.add_argument('--init-prune-epoch', type=int, default=0, help='The pruning rate / death rate.')
    parser.add_argument('--final-prune-epoch', type=int, default=1000, help='The density of the overall sparse network.')
    parser.add_argument('--fix', type=bool, default=False, help='Fix sparse connectivity during training. Default: True.')
    parser.add_argument('--sparse_init', type=str, default='uniform', help='sparse initialization: ERK, snip, Grasp')
    parser.add_argument('--growth', type=str, default='random', help='Growth mode. Choose from: momentum, random, random_unfired, and gradient.')
    parser.add_argument('--death', type=str, default='magnitude', help='Death mode / pruning mode. Choose from: magnitude, SET, threshold.')
    parser.add_argument('--redistribution', type=str, default='none', help='Redistribution mode. Choose from: momentum, magnitude, nonzeros, or none.')
    parser.add_argument('--death-rate', type=float, default=0.50, help='The pruning rate / death rate.')
    parser.add_argument('--density', type=float, default=0.3, help='The density of the overall sparse network.')
    parser.add_argument('--final_density', type=float, default=0.05, help='The density of the overall sparse network.')
    parser.add_argument('--update_frequency', type=int, default=5, metavar='N', help='how many iterations to train between parameter exploration')
    parser.add_argument('--decay-schedule', type=str, default='cosine', help='The decay schedule for the pruning rate. Default: cosine. Choose from: cosine, linear.')


class CosineDecay(object):
    """death-rate schedule; wraps torch's CosineAnnealingLR exactly like the reference."""

    def __init__(self, death_rate, T_max, eta_min=0.001, last_epoch=-1):
        self.sgd = optim.SGD(torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(1))]), lr=death_rate)
        self.cosine_stepper = torch.optim.lr_scheduler.CosineAnnealingLR(self.sgd, T_max, eta_min, last_epoch)

    def step(self):
        self.cosine_stepper.step()

    def get_dr(self):
        return self.sgd.param_groups[0]['lr']


class LinearDecay(object):
    def __init__(self, death_rate, factor=0.99, frequency=600):
        self.factor = factor
        self.steps = 0
        self.frequency = frequency

    def step(self):
        self.steps += 1

    def get_dr(self, death_rate):
        if self.steps > 0 and self.steps % self.frequency == 0:
            return death_rate * self.factor
        return death_rate


def _vp(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class Masking(object):
    def __init__(self, optimizer, death_rate=0.3, growth_death_ratio=1.0, death_rate_decay=None,
                 death_mode='magnitude', growth_mode='momentum', redistribution_mode='momentum', threshold=0.001,
                 train_loader=None, T_max=0., args=None):
        growth_modes = ['random', 'momentum', 'momentum_neuron', 'gradient']
        if growth_mode not in growth_modes:
            print('Growth mode: {0} not supported!'.format(growth_mode))
            print('Supported modes are:', str(growth_modes))
        self.args = args
        self.device = torch.device("cuda")
        self.growth_mode = growth_mode
        self.death_mode = death_mode
        self.growth_death_ratio = growth_death_ratio
        self.redistribution_mode = redistribution_mode
        self.death_rate_decay = death_rate_decay
        self.threshold = threshold

        self.masks = {}
        self.modules = []
        self.names = []
        self.optimizer = optimizer

        self.name2zeros = {}
        self.num_remove = {}
        self.num_death = {}
        self.name2nonzeros = {}
        self.death_rate = death_rate
        self.baseline_nonzero = None
        self.steps = 0
        self.explore_step = 0

        self.pruned_masks = {}
        self.regrowed_masks = {}
        self.pre_masks = None
        self.decay_flag = True

        self.total_nozeros = 0
        self.total_weights = 0
        self.loader = train_loader
        self.regrow_ratio = 1.01
        self.adv = self.args.adv
        self.curr_density = 0.0
        self.regrow_ones = 0
        self.T_max = T_max

        if self.args.fix:
            self.prune_every_k_steps = None
        else:
            self.prune_every_k_steps = self.args.update_frequency

        # fp32 association of the reference's nested sum(dim=-1) in kernel_death (core_channel.py:653-655).  The
        # reference's Masking lives on the GPU (it hard-codes .cuda()), where torch's reduce kernel computes
        # (a0 + a2) + a1 for three values; torch CPU folds left to right.  "cuda" reproduces the reference as it
        # actually runs; "cpu" reproduces a CPU run of it (the committed goldens were generated on CPU).
        self.sum_association = "cuda"
        self._tables = None       # cached device pointer tables for the multi-tensor kernels
        self._scratch = {}

    # ------------------------------------------------------------------ bookkeeping helpers
    def _assoc(self) -> int:
        if self.sum_association not in ("cuda", "cpu"):
            raise ValueError("Masking.sum_association must be 'cuda' or 'cpu'")
        return 1 if self.sum_association == "cuda" else 0

    def _params(self):
        """[(name, parameter)] of the masked tensors in named_parameters() order (= reference loop order)."""
        out = []
        for module in self.modules:
            for name, tensor in module.named_parameters():
                if name in self.masks:
                    out.append((name, tensor))
        return out

    def _check_cuda(self, t, what):
        if not t.is_cuda:
            raise _lib.E2EError("Masking (B200): %s lives on %s; the drop-in Masking runs on CUDA only "
                                "(the reference hard-codes .cuda() too, core_channel.py:67,326)" % (what, t.device))

    def _wire_modules(self):
        """hand every owning conv module its mask so that weight packing applies it on load."""
        for module in self.modules:
            owners = dict(module.named_modules())
            for name in self.masks:
                if not name.endswith(".weight"):
                    continue
                mod_name = name[:-len(".weight")]
                owner = owners.get(mod_name)
                if owner is None:
                    continue
                parent = owners.get(mod_name.rsplit(".", 1)[0]) if "." in mod_name else None
                if parent is not None and hasattr(parent, "e2e_weight_mask") and mod_name.endswith(".conv"):
                    parent.e2e_weight_mask = self.masks[name]
                else:
                    owner.e2e_weight_mask = self.masks[name]

    # ------------------------------------------------------------------ init (reference :109-287)
    def init(self, mode='ERK', density=0.05, erk_power_scale=1.0):
        self.density = density
        if mode == 'uniform':
            # reference :141-169 -- kernel-granular: one random.sample per tensor over row-major (C0, C1) pairs
            self.baseline_nonzero = 0
            for name, weight in self._params():
                density_n = 0.2 if weight.shape[0] == 48 else density
                k_size = np.prod(weight.shape[-3:])
                nonzeros = weight.numel() * density_n
                n_pairs = weight.shape[0] * weight.shape[1]
                kernel_num = round(nonzeros / k_size)
                idx_rand = random.sample(list(range(0, n_pairs)), kernel_num)
                pick = torch.as_tensor(idx_rand, dtype=torch.int64)
                m2 = torch.zeros(n_pairs, dtype=torch.float32)
                m2[pick] = 1.0
                m2 = m2.view(weight.shape[0], weight.shape[1], 1, 1, 1).to(self.masks[name].device)
                self.masks[name].copy_(m2.expand_as(self.masks[name]))
                nnz = kernel_num * int(k_size)
                self.baseline_nonzero += nnz
                print(f"layer: {name}, shape: {self.masks[name].shape}, density: {nnz / self.masks[name].numel()}")
        elif mode == 'GMP':
            self.baseline_nonzero = 0
            for name, weight in self._params():
                self.masks[name].fill_(1.0)
                self.baseline_nonzero += self.masks[name].numel()
        elif mode == 'uniform_ori':
            self.baseline_nonzero = 0
            for name, weight in self._params():
                self.masks[name][:] = (torch.rand(weight.shape) < density).float().data.to(self.masks[name].device)
                self.baseline_nonzero += weight.numel() * density
                print(f"layer: {name}, shape: {self.masks[name].shape}, density: {density}")
        elif mode == 'lottery_ticket':
            print('initialize by lottery ticket')
            self.baseline_nonzero = 0
            scores = torch.cat([torch.abs(w).flatten() for _, w in self._params()])
            keep = int(len(scores) * self.density)
            thr, _ = torch.topk(scores, keep, sorted=True)
            for name, weight in self._params():
                self.masks[name].copy_((torch.abs(weight) >= thr[-1]).float())
                self.baseline_nonzero += int((self.masks[name] != 0).sum().item())
        elif mode == 'ERK':
            self._init_erk(erk_power_scale)
        else:
            raise NotImplementedError("sparse_init=%r is not on the E2ENet hot path (reference default: 'uniform'; "
                                      "snip/GraSP need a data loader and are unreachable from simple_main.py)" % mode)

        self._wire_modules()
        self.apply_mask()
        self.fired_masks = copy.deepcopy(self.masks)   # used for ITOP
        total_size = sum(w.numel() for w in self.masks.values())
        print('Total Model parameters:', total_size)
        sparse_size = sum(int((w != 0).sum().item()) for w in self.masks.values())
        print('Total parameters under sparsity level of {0}: {1}'.format(self.density, sparse_size / total_size))

    def _init_erk(self, erk_power_scale):
        # element-wise Erdos-Renyi-Kernel init of the reference (:202-272); legacy mode, plain torch
        print('initialize by ERK')
        total_params = sum(w.numel() for w in self.masks.values())
        dense_layers = set()
        while True:
            divisor, rhs, raw = 0, 0, {}
            for name, mask in self.masks.items():
                n_param = np.prod(mask.shape)
                if name in dense_layers:
                    rhs -= n_param * (1 - self.density)
                else:
                    rhs += n_param * self.density
                    raw[name] = (np.sum(mask.shape) / np.prod(mask.shape)) ** erk_power_scale
                    divisor += raw[name] * n_param
            epsilon = rhs / divisor
            max_prob = np.max(list(raw.values()))
            if max_prob * epsilon > 1:
                for k, v in raw.items():
                    if v == max_prob:
                        print(f"Sparsity of var:{k} had to be set to 0.")
                        dense_layers.add(k)
            else:
                break
        total_nonzero = 0.0
        for name, mask in self.masks.items():
            d = 1.0 if name in dense_layers else epsilon * raw[name]
            print(f"layer: {name}, shape: {mask.shape}, density: {d}")
            self.masks[name][:] = (torch.rand(mask.shape) < d).float().data.to(mask.device)
            total_nonzero += d * mask.numel()
        print(f"Overall sparsity {total_nonzero / total_params}")

    # ------------------------------------------------------------------ step (reference :290-317)
    def step(self, _mask_already_applied=False):
        """reference :290-317.  `_mask_already_applied` (extension, default off): the apply_mask of this
        iteration already ran on the device -- it is part of a replayed CUDA graph of the training step
        (training.TrainStep.enable_graph) -- so only the host bookkeeping and, on update iterations,
        the prune / regrow run here."""
        if not _mask_already_applied:
            self.apply_mask()
        if self.decay_flag:
            self.death_rate_decay.step()
            self.death_rate = self.death_rate_decay.get_dr()
        else:
            self.death_rate = 0.001
            self.adv = False
        self.steps += 1
        if self.prune_every_k_steps is not None:
            if self.steps % self.prune_every_k_steps == 0:
                self.explore_step += 1
                self.truncate_weights()
                self.cal_nonzero_counts()
                self.curr_density = self.total_nozeros / self.total_weights
                print('curr_density: {0:.4f}, final_density:{1:.4f}'.format(self.curr_density, self.args.final_density))
                _, _ = self.fired_masks_update()
                if self.explore_step > 1:
                    self.print_nonzero_counts()
                self.pre_masks = copy.deepcopy(self.pruned_masks)

    def add_module(self, module, density, sparse_init='ER'):
        self.modules.append(module)
        self.module = module
        for name, tensor in module.named_parameters():
            if ('loc' in name and 'context' not in name) or 'up' in name:     # reference :324
                self._check_cuda(tensor, "parameter " + name)
                self.names.append(name)
                self.masks[name] = torch.zeros_like(tensor, dtype=torch.float32, requires_grad=False)
        print('Removing biases...')
        self.remove_weight_partial_name('bias')
        print('Removing biases...')
        self.remove_weight_partial_name('instnorm')
        print('Removing 2D batch norms...')
        self.remove_type(nn.BatchNorm2d)
        print('Removing 1D batch norms...')
        self.remove_type(nn.BatchNorm1d)
        self.init(mode=sparse_init, density=density)

    # ------------------------------------------------------------------ counters
    def _counts_all(self, update_fired=False):
        """(nnz, fired_nnz) per masked tensor with a single device->host read."""
        lib = _lib.load()
        items = list(self.masks.items())
        buf = torch.empty(len(items) * 2, dtype=torch.int32, device=items[0][1].device)
        for i, (name, m) in enumerate(items):
            fired = None
            if update_fired:
                f = self.fired_masks[name]
                if f.dtype != torch.uint8:
                    f = f.to(torch.uint8)
                    self.fired_masks[name] = f
                fired = f
            _lib.check(lib.e2e_mask_counts(_vp(m), _vp(fired), m.numel(),
                                           C.c_void_p(buf.data_ptr() + 8 * i), _lib.stream_ptr()), "mask_counts")
        host = buf.cpu().numpy().reshape(-1, 2)
        return {name: (int(host[i, 0]), int(host[i, 1])) for i, (name, _) in enumerate(items)}

    def cal_nonzero_counts(self):
        counts = self._counts_all()
        self.total_nozeros = 0
        self.total_weights = 0
        for name, _ in self._params():
            self.total_nozeros += counts[name][0]
            self.total_weights += self.masks[name].numel()

    def remove_weight(self, name):
        if name in self.masks:
            print('Removing {0} of size {1} = {2} parameters.'.format(name, self.masks[name].shape, self.masks[name].numel()))
            self.masks.pop(name)
        elif name + '.weight' in self.masks:
            print('Removing {0} of size {1} = {2} parameters.'.format(name, self.masks[name + '.weight'].shape,
                                                                      self.masks[name + '.weight'].numel()))
            self.masks.pop(name + '.weight')
        else:
            print('ERROR', name)

    def remove_weight_partial_name(self, partial_name):
        removed = set()
        for name in list(self.masks.keys()):
            if partial_name in name:
                print('Removing {0} of size {1} with {2} parameters...'.format(name, self.masks[name].shape,
                                                                               np.prod(self.masks[name].shape)))
                removed.add(name)
                self.masks.pop(name)
        print('Removed {0} layers.'.format(len(removed)))
        self.names = [n for n in self.names if n not in removed]
        self._tables = None

    def remove_type(self, nn_type):
        for module in self.modules:
            for name, module in module.named_modules():
                if isinstance(module, nn_type):
                    self.remove_weight(name)

    # ------------------------------------------------------------------ apply_mask (reference :427-434)
    def apply_mask(self):
        """w *= mask and momentum_buffer *= mask for every masked tensor: one multi-tensor kernel."""
        params = self._params()
        if not params:
            return
        lib = _lib.load()
        from .. import ops
        ops.bump_weight_epoch()          # weights / masks change below through raw pointers
        moms = []
        for name, p in params:
            self._check_cuda(p, "parameter " + name)
            st = self.optimizer.state[p] if p in self.optimizer.state else {}
            moms.append(st.get('momentum_buffer', None) if isinstance(st, dict) else None)
        for has_mom in (True, False):
            sel = [(n, p, b) for (n, p), b in zip(params, moms) if (b is not None) == has_mom]
            if not sel:
                continue
            key = (has_mom,) + tuple((p.data_ptr(), self.masks[n].data_ptr(), b.data_ptr() if b is not None else 0)
                                     for n, p, b in sel)
            tab = self._scratch.get(("apply", has_mom))
            if tab is None or tab[0] != key:
                dev = sel[0][1].device
                for n, p, b in sel:
                    assert p.dtype == torch.float32 and p.is_contiguous() and self.masks[n].is_contiguous()
                    assert b is None or (b.dtype == torch.float32 and b.is_contiguous())
                mk = lambda vals: torch.tensor(vals, dtype=torch.int64).to(dev)
                wt = mk([p.data_ptr() for _, p, _ in sel])
                mt = mk([self.masks[n].data_ptr() for n, _, _ in sel])
                bt = mk([b.data_ptr() for _, _, b in sel]) if has_mom else None
                nt = mk([p.numel() for _, p, _ in sel])
                tab = (key, wt, bt, mt, nt, max(p.numel() for _, p, _ in sel))
                self._scratch[("apply", has_mom)] = tab
            _, wt, bt, mt, nt, mx = tab
            _lib.check(lib.e2e_mask_apply_multi(_vp(wt), _vp(bt), _vp(mt), _vp(nt), len(sel), mx, _lib.stream_ptr()),
                       "mask_apply_multi")

    # ------------------------------------------------------------------ prune / regrow (reference :556-611)
    def truncate_weights(self):
        if self.death_mode != 'magnitude':
            raise NotImplementedError("death_mode=%r: only 'magnitude' (kernel_death) is wired to the kernel-granular "
                                      "path in the reference (core_channel.py:566-574)" % self.death_mode)
        lib = _lib.load()
        params = self._params()
        dev = params[0][1].device
        counts = self._counts_all()
        n = len(params)
        kill_counts = torch.zeros(n * 2, dtype=torch.int32, device=dev)
        dead_lists = {}
        # ---- death for ALL tensors first
        for i, (name, weight) in enumerate(params):
            mask = self.masks[name]
            self.name2nonzeros[name] = float(counts[name][0])
            self.name2zeros[name] = mask.numel() - self.name2nonzeros[name]
            k_size = int(np.prod(weight.shape[-3:]))
            n_kernels = weight.shape[0] * weight.shape[1]
            prune_num = math.ceil(self.death_rate * self.name2nonzeros[name] / k_size)
            num_zeros = math.ceil(self.name2zeros[name] / k_size)
            rank = num_zeros + prune_num - 1
            if rank >= n_kernels:
                raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (rank, n_kernels))
            if rank < 0:
                rank += n_kernels          # value[-1]: python negative indexing of the sorted vector
            l1 = torch.empty(n_kernels, dtype=torch.float32, device=dev)
            thr = torch.empty(1, dtype=torch.float32, device=dev)
            kd, kh, kw = (int(s) for s in weight.shape[-3:])
            w = weight.data
            assert w.is_contiguous() and w.dtype == torch.float32
            _lib.check(lib.e2e_mask_kernel_l1(_vp(w), n_kernels, kd, kh, kw, self._assoc(), _vp(l1), _lib.stream_ptr()),
                       "kernel_l1")
            _lib.check(lib.e2e_mask_kth(_vp(l1), n_kernels, rank, _vp(thr), None, _lib.stream_ptr()), "mask_kth")
            _lib.check(lib.e2e_mask_kill(_vp(l1), _vp(thr), _vp(mask), n_kernels, k_size,
                                         C.c_void_p(kill_counts.data_ptr() + 8 * i), _lib.stream_ptr()), "mask_kill")
            self.num_death[name] = prune_num
            dead = torch.empty(n_kernels, dtype=torch.int32, device=dev)
            nd = torch.empty(1, dtype=torch.int32, device=dev)
            _lib.check(lib.e2e_mask_dead_list(_vp(mask), n_kernels, k_size, _vp(dead), _vp(nd), None,
                                              _lib.stream_ptr()), "mask_dead_list")
            dead_lists[name] = dead
        kc = kill_counts.cpu().numpy().reshape(-1, 2)        # the one read between death and growth
        for i, (name, weight) in enumerate(params):
            k_size = int(np.prod(weight.shape[-3:]))
            self.num_remove[name] = int(self.name2nonzeros[name] - int(kc[i, 0]) * k_size)
            self.pruned_masks[name] = self.masks[name].clone()
        # ---- growth for ALL tensors (python RNG on the host, same call sequence as the reference)
        for i, (name, weight) in enumerate(params):
            if self.growth_mode == 'random':
                n_dead = int(kc[i, 1])
                k_size = int(np.prod(weight.shape[-3:]))
                idx_rand = random.sample(list(range(0, n_dead)), self.num_death[name])
                if idx_rand:
                    pick = torch.tensor(idx_rand, dtype=torch.int32).to(dev, non_blocking=True)
                    _lib.check(lib.e2e_mask_grow(_vp(self.masks[name]), _vp(dead_lists[name]), _vp(pick),
                                                 len(idx_rand), k_size, _lib.stream_ptr()), "mask_grow")
            elif self.growth_mode == 'gradient':
                new_mask = self.kernel_grad_growth(name, self.masks[name].data.byte(), weight)
                self.masks[name].copy_(new_mask.float())
            else:
                raise NotImplementedError("growth_mode=%r: only 'random' (kernel_growth) and 'gradient' "
                                          "(kernel_grad_growth) act on kernels in the reference" % self.growth_mode)
            self.regrowed_masks[name] = self.masks[name]
        self.apply_mask()

    # -- reference-signature entry points (standalone use; they sync like the reference does)
    def kernel_death(self, mask, weight, name):
        from .. import ops
        ops.bump_weight_epoch()
        lib = _lib.load()
        k_size = int(np.prod(weight.shape[-3:]))
        n_kernels = weight.shape[0] * weight.shape[1]
        prune_num = math.ceil(self.death_rate * self.name2nonzeros[name] / k_size)
        num_zeros = math.ceil(self.name2zeros[name] / k_size)
        rank = num_zeros + prune_num - 1
        if rank >= n_kernels:
            raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (rank, n_kernels))
        if rank < 0:
            rank += n_kernels
        dev = weight.device
        l1 = torch.empty(n_kernels, dtype=torch.float32, device=dev)
        thr = torch.empty(1, dtype=torch.float32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)
        kd, kh, kw = (int(s) for s in weight.shape[-3:])
        _lib.check(lib.e2e_mask_kernel_l1(_vp(weight.data), n_kernels, kd, kh, kw, self._assoc(), _vp(l1), _lib.stream_ptr()),
                   "kernel_l1")
        _lib.check(lib.e2e_mask_kth(_vp(l1), n_kernels, rank, _vp(thr), None, _lib.stream_ptr()), "mask_kth")
        _lib.check(lib.e2e_mask_kill(_vp(l1), _vp(thr), _vp(mask), n_kernels, k_size, _vp(cnt), _lib.stream_ptr()), "mask_kill")
        return mask, prune_num

    def kernel_growth(self, name, new_mask, weight):
        from .. import ops
        ops.bump_weight_epoch()
        lib = _lib.load()
        num_growth = self.num_death[name]
        out = new_mask.float().contiguous().clone()
        k_size = int(np.prod(weight.shape[-3:]))
        n_kernels = weight.shape[0] * weight.shape[1]
        dead = torch.empty(n_kernels, dtype=torch.int32, device=out.device)
        nd = torch.empty(1, dtype=torch.int32, device=out.device)
        _lib.check(lib.e2e_mask_dead_list(_vp(out), n_kernels, k_size, _vp(dead), _vp(nd), None, _lib.stream_ptr()),
                   "mask_dead_list")
        idx_rand = random.sample(list(range(0, int(nd.item()))), num_growth)
        if idx_rand:
            pick = torch.tensor(idx_rand, dtype=torch.int32, device=out.device)
            _lib.check(lib.e2e_mask_grow(_vp(out), _vp(dead), _vp(pick), len(idx_rand), k_size, _lib.stream_ptr()), "mask_grow")
        return out.to(new_mask.dtype)

    def kernel_grad_growth(self, name, new_mask, weight):
        from .. import ops
        ops.bump_weight_epoch()
        # reference :771-790 (growth='gradient'); bookkeeping-sized torch ops on (C0, C1) matrices
        num_growth = self.num_death[name]
        if num_growth == 0:
            return new_mask
        mask_sum = torch.squeeze(torch.sum(torch.sum(torch.abs(new_mask), dim=-1), dim=-1))
        data_sum = torch.squeeze(torch.sum(torch.sum(torch.abs(self.get_gradient_for_weights(weight)), dim=-1), dim=-1))
        grad = data_sum * (mask_sum < 1).float()
        value, _ = torch.sort(grad.data.view(-1), descending=True)
        idx = torch.nonzero(grad.data > value[num_growth].item())
        new_mask[idx[:, 0], idx[:, 1]] = 1.0
        return new_mask

    # ------------------------------------------------------------------ utility (reference :824-881)
    def get_momentum_for_weight(self, weight):
        if 'exp_avg' in self.optimizer.state[weight]:
            adam_m1 = self.optimizer.state[weight]['exp_avg']
            adam_m2 = self.optimizer.state[weight]['exp_avg_sq']
            grad = adam_m1 / (torch.sqrt(adam_m2) + 1e-08)
        elif 'momentum_buffer' in self.optimizer.state[weight]:
            grad = self.optimizer.state[weight]['momentum_buffer']
        return grad

    def get_gradient_for_weights(self, weight):
        return weight.grad.clone()

    def print_nonzero_counts(self):
        counts = self._counts_all()
        for name, tensor in self._params():
            mask = self.masks[name]
            num_nonzeros = counts[name][0]
            a = self.pre_masks[name].data >= 1.0
            b = self.pruned_masks[name].data >= 1.0
            diff = int((a != b).sum().item())
            print('{0}: {1}->{2}, density: {3:.3f}, diff: {4}'.format(name, self.name2nonzeros[name], num_nonzeros,
                                                                      num_nonzeros / float(mask.numel()), diff))
        for name, tensor in self._params():
            print('Death rate: {0}\n'.format(self.death_rate))
            break

    def fired_masks_update(self):
        counts = self._counts_all(update_fired=True)
        ntotal_fired_weights = 0.0
        ntotal_weights = 0.0
        layer_fired_weights = {}
        for name, weight in self._params():
            fired = float(counts[name][1])
            numel = float(self.fired_masks[name].numel())
            ntotal_fired_weights += fired
            ntotal_weights += numel
            layer_fired_weights[name] = fired / numel
            print('Layerwise percentage of the fired weights of', name, 'is:', layer_fired_weights[name])
        total_fired_weights = ntotal_fired_weights / ntotal_weights
        print('The percentage of the total fired weights is:', total_fired_weights)
        return layer_fired_weights, total_fired_weights

    def death_decay_update(self, death_rate_decay=None, decay_flag=True):
        self.death_rate_decay = death_rate_decay
        self.decay_flag = decay_flag
