"""Device-resident replay buffers -- B200-native drop-in for the reference's my_replay_buffer
(my_replay_buffer.py:6-128).

Storage is ONE fp32 array-of-rows tensor in HBM (a transition = one contiguous, 32-byte-aligned
row), so ``add`` is a single small H2D copy and ``sample`` is one coalesced gather kernel; the
reference keeps five/seven float64 NumPy arrays on the host and casts + uploads at every sample.
Storing fp32 at ``add`` is bit-identical to the reference's ``torch.FloatTensor(float64[ind])`` at
``sample`` (both are one round-to-nearest-even double->float conversion).

Same public surface: ``__init__(obs_space, action_space, max_size=int(1e6), load_folder=None)``,
``add``, ``sample``, ``save``, ``load``, attributes ``ptr size max_size device store_np store_pkl``
(the per-field arrays are available as float64 NumPy snapshots).  Extras: ``add_batch`` (bulk
ingest) and ``sample(batch, indices=...)`` / ``rng="device"``.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle

import numpy as np
import torch

from . import _lib

_STAGE_SLOTS = 256


class _DeviceReplay(object):
    # subclasses define: fields (name -> (offset, shape)), store_np order
    def _init_storage(self, max_size, load_folder, n_agents=1):
        self._lib = _lib.require_cuda()
        self.max_size = int(max_size)
        self.n_agents = int(n_agents)
        self.store_pkl = ["ptr", "size"]
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.row_floats = max(off + int(np.prod(shape)) for off, shape in self._fields.values())
        self.row_stride = (self.row_floats + 7) // 8 * 8          # rows start on 32-byte sector boundaries
        # population (n_agents > 1): one ring per member, a constant stride apart; self._rows is member 0's ring
        self._all_rows = torch.zeros(self.n_agents, self.max_size, self.row_stride, dtype=torch.float32, device=self.device)
        self._rows = self._all_rows[0]
        self._ptrs, self._sizes = [0] * self.n_agents, [0] * self.n_agents
        self._stage = torch.zeros(_STAGE_SLOTS, self.row_floats, dtype=torch.float32).pin_memory()
        self._stage_np = self._stage.numpy()
        self._rows_ptr, self._stage_ptr = self._rows.data_ptr(), self._stage.data_ptr()
        self._slot = 0
        self.ptr = 0
        self.size = 0
        self.rng = "host"
        self._philox_step = 0
        if load_folder is not None:
            self.load(load_folder)

    # ------------------------------------------------------------------ C-ABI view
    def _view(self) -> _lib.ReplayView:
        v = self.__dict__.get("_cached_view")
        if v is None:                                        # the storage never moves: only `size` changes between calls
            v = _lib.ReplayView()
            v.rows = C.c_void_p(self._rows.data_ptr())
            v.row_stride, v.row_floats = self.row_stride, self.row_floats
            v.max_size, v.agent_stride = self.max_size, (self.max_size * self.row_stride if self.n_agents > 1 else 0)
            self.__dict__["_cached_view"] = v
        v.size = self.size if self.n_agents == 1 else min(self.size, *self._sizes[1:])   # lock-step: common valid prefix
        return v

    # ------------------------------------------------------------------ add (my_replay_buffer.py:46-56,109-117)
    def _next_slot(self):
        if self._slot == _STAGE_SLOTS:
            torch.cuda.current_stream().synchronize()       # the ring of pinned slots is about to be reused
            self._slot = 0
        s = self._slot
        self._slot += 1
        return s

    def _commit_row(self, slot, agent=0):
        if agent:                                            # population member >= 1: its own ring, pointer and size
            ring = self._rows_ptr + agent * self.max_size * self.row_stride * 4
            rc = self._lib.rb_add_rows(ring, self.row_stride, self.row_floats, self.max_size, self._ptrs[agent],
                                       self._stage_ptr + slot * self.row_floats * 4, 1, _lib.stream_ptr())
            if rc:
                _lib.check(rc)
            self._ptrs[agent] = (self._ptrs[agent] + 1) % self.max_size
            self._sizes[agent] = min(self._sizes[agent] + 1, self.max_size)
            return
        rc = self._lib.rb_add_rows(self._rows_ptr, self.row_stride, self.row_floats, self.max_size, self.ptr,
                                   self._stage_ptr + slot * self.row_floats * 4, 1, _lib.stream_ptr())
        if rc:
            _lib.check(rc)
        self.ptr = (self.ptr + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def add_batch(self, agent=0, **columns):
        """Bulk ``add`` of n transitions given as arrays keyed by field name (``done`` instead of
        ``not_done``); equivalent to n calls of ``add`` in order, including ring wrap-around.
        ``agent`` selects the population member whose ring receives them."""
        if agent:
            return self._add_batch_member(int(agent), columns)
        n = len(next(iter(columns.values())))
        chunk = max(1, min(n, (64 << 20) // (4 * self.row_floats)))
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            host = np.empty((hi - lo, self.row_floats), dtype=np.float32)
            for name, (off, shape) in self._fields.items():
                w = int(np.prod(shape))
                if name == "not_done":
                    src = 1.0 - np.asarray(columns["done"][lo:hi], dtype=np.float64).reshape(hi - lo, w)
                else:
                    src = np.asarray(columns[name][lo:hi]).reshape(hi - lo, w)
                host[:, off:off + w] = src
            pinned = torch.from_numpy(host).pin_memory()
            done = 0
            while done < hi - lo:                           # rb_add_rows wraps once per call
                m = min(hi - lo - done, self.max_size)
                _lib.check(self._lib.rb_add_rows(C.c_void_p(self._rows.data_ptr()), self.row_stride, self.row_floats,
                                                 self.max_size, self.ptr, C.c_void_p(pinned[done].data_ptr()), m,
                                                 _lib.stream_ptr()))
                self.ptr = (self.ptr + m) % self.max_size
                self.size = min(self.size + m, self.max_size)
                done += m
            torch.cuda.current_stream().synchronize()       # pinned chunk is released after this

    def _add_batch_member(self, agent, columns):
        """add_batch for population member ``agent`` >= 1 (member 0 keeps the reference attributes ptr/size)."""
        n = len(next(iter(columns.values())))
        host = np.empty((n, self.row_floats), dtype=np.float32)
        for name, (off, shape) in self._fields.items():
            w = int(np.prod(shape))
            src = (1.0 - np.asarray(columns["done"], dtype=np.float64).reshape(n, w)) if name == "not_done" \
                else np.asarray(columns[name]).reshape(n, w)
            host[:, off:off + w] = src
        rows = self._all_rows[agent]
        idx = (self._ptrs[agent] + np.arange(n)) % self.max_size
        rows[torch.as_tensor(idx, device=self.device), :self.row_floats] = torch.from_numpy(host).to(self.device)
        self._ptrs[agent] = int((self._ptrs[agent] + n) % self.max_size)
        self._sizes[agent] = min(self._sizes[agent] + n, self.max_size)

    # ------------------------------------------------------------------ sample (:58-69,119-128)
    def sample(self, batch_size, indices=None):
        """Uniform sampling with replacement; returns float32 device tensors in the reference's order.
        ``indices`` (host ints or a device int64 tensor) overrides the draw.  With ``self.rng ==
        "host"`` (default) the draw is the reference's ``np.random.randint(0, size, batch)`` from the
        global NumPy generator; ``"device"`` uses Philox on the GPU and never touches the host."""
        batch_size = int(batch_size)
        s = _lib.stream_ptr()
        if indices is None and self.rng == "device":
            if self.size <= 0:
                raise ValueError("high <= 0")
            idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
            _lib.check(self._lib.rb_philox_indices(C.c_void_p(idx.data_ptr()), batch_size, self.size,
                                                   getattr(self, "seed", 0), 2, self._philox_step, s))
            self._philox_step += 1
        else:
            if indices is None:
                indices = np.random.randint(0, self.size, size=batch_size)       # raises ValueError when empty
            if isinstance(indices, torch.Tensor):
                idx = indices.to(self.device, torch.int64).contiguous()
            else:
                idx = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=self.device)
        outs = [torch.empty((batch_size, *shape), dtype=torch.float32, device=self.device)
                for _, shape in self._fields.values()]
        n = len(outs)
        view = self._view()
        seg_off = (C.c_int64 * n)(*[off for off, _ in self._fields.values()])
        seg_len = (C.c_int64 * n)(*[int(np.prod(shape)) for _, shape in self._fields.values()])
        dst = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        dst_ld = (C.c_int64 * n)(*[int(np.prod(shape)) for _, shape in self._fields.values()])
        _lib.check(self._lib.rb_sample_indices(C.byref(view), C.c_void_p(idx.data_ptr()), batch_size, n, seg_off,
                                               seg_len, dst, dst_ld, s))
        return tuple(outs)

    # ------------------------------------------------------------------ persistence (:28-44,91-107)
    def _field_array(self, name) -> np.ndarray:
        off, shape = self._fields[name]
        w = int(np.prod(shape))
        return self._rows[:, off:off + w].cpu().numpy().astype(np.float64).reshape(self.max_size, *shape)

    def save(self, folder):
        os.makedirs(folder, exist_ok=True)
        for attrib in self.store_pkl:
            with open(os.path.join(folder, attrib + ".pkl"), "wb") as f:
                pickle.dump(getattr(self, attrib), f, protocol=4)
        for attrib in self.store_np:                         # np.save payloads in files named *.pkl, as the reference
            with open(os.path.join(folder, attrib + ".pkl"), "wb") as f:
                np.save(f, self._field_array(attrib))

    def load(self, folder):
        for attrib in self.store_pkl:
            with open(os.path.join(folder, attrib + ".pkl"), "rb") as f:
                setattr(self, attrib, int(pickle.load(f)))
        arrays = {}
        for attrib in self.store_np:
            with open(os.path.join(folder, attrib + ".pkl"), "rb") as f:
                arrays[attrib] = np.load(f)
        n = int(next(iter(arrays.values())).shape[0])
        if n != self.max_size:
            # the reference adopts the stored arrays whatever max_size the constructor was given (my_replay_buffer.py:38-44:
            # experience_injection.py builds the buffer with the default max_size and loads a smaller one): so do we
            if self.n_agents != 1:
                raise ValueError(f"stored buffer has {n} rows, this population's rings have {self.max_size}")
            torch.cuda.current_stream().synchronize()
            self.max_size = n
            self._all_rows = torch.zeros(1, n, self.row_stride, dtype=torch.float32, device=self.device)
            self._rows = self._all_rows[0]
            self._rows_ptr = self._rows.data_ptr()
            self.__dict__.pop("_cached_view", None)
        for attrib, arr in arrays.items():
            off, shape = self._fields[attrib]
            w = int(np.prod(shape))
            self._rows[:, off:off + w] = torch.from_numpy(arr.reshape(self.max_size, w).astype(np.float32)).to(self.device)

    def get_rows(self, name, start=0, stop=None) -> np.ndarray:
        """float64 copy of rows [start, stop) of one field (``rb.state`` & co. snapshot the WHOLE ring on every access and are
        read-only: an in-place edit of such a snapshot does not reach the device)."""
        off, shape = self._fields[name]
        w = int(np.prod(shape))
        stop = self.max_size if stop is None else stop
        return self._rows[start:stop, off:off + w].cpu().numpy().astype(np.float64).reshape(stop - start, *shape)

    def set_rows(self, name, start, values):
        """Write rows of one field in place (what ``rb.reward[i] = x`` does on the reference's NumPy arrays)."""
        off, shape = self._fields[name]
        w = int(np.prod(shape))
        v = np.asarray(values, dtype=np.float64).reshape(-1, w).astype(np.float32)
        self._rows[start:start + v.shape[0], off:off + w] = torch.from_numpy(v).to(self.device)

    def __getattr__(self, name):
        # float64 NumPy snapshots of the per-field arrays (reference attribute names)
        fields = self.__dict__.get("_fields")
        if fields is not None and name in fields:
            return self._field_array(name)
        raise AttributeError(name)


def _layout(spec):
    fields, off = {}, 0
    for name, shape in spec:
        fields[name] = (off, tuple(shape))
        off += int(np.prod(shape))
    return fields


class ReplayBuffer_featured(_DeviceReplay):
    """Row = [state | action | next_state | reward | not_done]  (my_replay_buffer.py:72-128)."""

    def __init__(self, obs_space, action_space, max_size=int(1e6), load_folder=None, n_agents=1):
        S, A = obs_space.shape[0], action_space.shape[0]
        self.store_np = ["state", "action", "next_state", "reward", "not_done"]
        self._fields = _layout([("state", (S,)), ("action", (A,)), ("next_state", (S,)), ("reward", (1,)), ("not_done", (1,))])
        self._init_storage(max_size, load_folder, n_agents)

    def add(self, state, action, next_state, reward, done, agent=0):
        slot = self._next_slot()
        row, f = self._stage_np[slot], self._fields
        S, A = f["state"][1][0], f["action"][1][0]
        row[0:S] = state
        row[S:S + A] = action
        row[S + A:2 * S + A] = next_state
        row[2 * S + A] = reward
        row[2 * S + A + 1] = 1. - done
        self._commit_row(slot, agent)


class ReplayBuffer_particles(_DeviceReplay):
    """Fields of my_replay_buffer.py:6-69.  Row = [particles | next_particles | features | action | next_features | reward |
    not_done]: the two particle sets (N*D floats each, padded to a multiple of 4) lead the row so that both start on a
    16-byte boundary -- what the gather kernel's bulk-copy (cp.async.bulk) staging of large rows needs
    (csrc/misc.cuh: gather_row)."""

    def __init__(self, obs_space, action_space, max_size=int(1e6), load_folder=None, n_agents=1):
        F, pshape, A = obs_space[0].shape[0], tuple(obs_space[1].shape), action_space.shape[0]
        self.store_np = ["state_features", "state_particles", "action", "next_state_features", "next_state_particles",
                         "reward", "not_done"]
        pn = int(np.prod(pshape))
        pnp = (pn + 3) // 4 * 4
        o_f = 2 * pnp
        # dict order = the reference's sample() order (:61-69); offsets = the device row layout above
        self._fields = {"state_features": (o_f, (F,)), "state_particles": (0, pshape), "action": (o_f + F, (A,)),
                        "next_state_features": (o_f + F + A, (F,)), "next_state_particles": (pnp, pshape),
                        "reward": (o_f + 2 * F + A, (1,)), "not_done": (o_f + 2 * F + A + 1, (1,))}
        self._init_storage(max_size, load_folder, n_agents)

    def add(self, state, action, next_state, reward, done, agent=0):
        slot = self._next_slot()
        row = self._stage_np[slot]
        vals = (state[0], state[1], action, next_state[0], next_state[1], reward, 1. - done)
        for (off, shape), v in zip(self._fields.values(), vals):
            w = int(np.prod(shape))
            row[off:off + w] = np.asarray(v).reshape(-1) if w > 1 else v
        self._commit_row(slot, agent)


# experience_injection.py:3 imports a name the reference never defines; its environment
# has tuple observations, so the particles buffer is what it means.
ReplayBuffer = ReplayBuffer_particles
