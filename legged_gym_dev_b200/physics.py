"""Physics stand-ins behind the LeggedRobot drop-in.

The reference talks to PhysX through `self.gym` and three aliased state tensors
(legged_robot.py:92-96,111-112,537-551).  PhysX is closed source and out of scope, so the env takes a
`physics` object with the same three tensors and the same three moments of the step:

    simulate(torques)   <- gym.set_dof_actuation_force_tensor + gym.simulate + gym.refresh_dof_state_tensor
    refresh()           <- gym.refresh_actor_root_state_tensor + gym.refresh_net_contact_force_tensor
    commit_resets()     <- gym.set_dof_state_tensor_indexed / set_actor_root_state_tensor_indexed (:428-430,452-454)

`ReplayPhysics` replays a recorded / synthetic tape (legged_gym_dev_b200.synthetic).  An Isaac Gym
adapter implementing the same three calls is sketched in INTEGRATION.md.
"""
import torch


class ReplayPhysics:
    """copy=True : frames are copied into persistent aliased buffers (what PhysX does; resets written by
                   the kernels are overwritten by the next frame).  Used by parity tests.
       copy=False: the aliased tensors are re-pointed at the tape frames (zero-copy replay; in-kernel
                   resets then land in the tape).  Used by the throughput bench so that no
                   device-to-device copy sits inside the timed region (SURVEY.md §8d cfg 2)."""

    def __init__(self, tape, device="cuda", copy=True):
        self.tape = tape
        self.copy = copy
        dev = torch.device(device)
        self.root_frames = tape.root.to(dev)
        self.dof_frames = tape.dof.to(dev)
        self.contact_frames = tape.contact.to(dev).reshape(tape.frames, tape.num_envs * tape.contact.shape[2], 3)
        self.num_envs, self.frames, self.decimation = tape.num_envs, tape.frames, tape.decimation
        self.frame = 0
        self.sub = 0
        # per-frame views are built once: indexing a tensor costs microseconds of Python per call
        self._root_v = [self.root_frames[f] for f in range(self.frames)]
        self._contact_v = [self.contact_frames[f] for f in range(self.frames)]
        self._dof_v = [[self.dof_frames[f, j] for j in range(self.decimation)] for f in range(self.frames)]
        if copy:
            self.root_states = self.root_frames[0].clone()
            self.dof_state = self.dof_frames[0, 0].clone()
            self.contact_forces = self.contact_frames[0].clone()
        else:
            self.root_states = self._root_v[0]
            self.dof_state = self._dof_v[0][0]
            self.contact_forces = self._contact_v[0]

    def simulate(self, torques):
        v = self._dof_v[self.frame % self.frames][self.sub]
        if self.copy:
            self.dof_state.copy_(v)
        else:
            self.dof_state = v
        self.sub += 1

    def refresh(self):
        f = self.frame % self.frames
        if self.copy:
            self.root_states.copy_(self._root_v[f])
            self.contact_forces.copy_(self._contact_v[f])
        else:
            self.root_states = self._root_v[f]
            self.contact_forces = self._contact_v[f]
        self.frame += 1
        self.sub = 0

    def commit_resets(self, reset_buf):
        """PhysX would take over the reset rows of dof_state / root_states here; a replay has nothing to do."""
        return None


class HostReplayPhysics(ReplayPhysics):
    """Frames live in PINNED HOST memory and cross PCIe inside every step: the end-to-end arm of bench.py (a physics
    engine that hands its state over from the host every sub-step).  The copies of step s+1 are issued on a side
    stream into a second set of device buffers while step s computes (double buffering, like any input pipeline);
    every byte still moves inside the timed region."""

    def __init__(self, tape, device="cuda", return_torques=True):
        self.tape, self.copy = tape, True
        dev = self.device = torch.device(device)
        self.return_torques = return_torques
        self.root_frames = tape.root.pin_memory()
        self.dof_frames = tape.dof.pin_memory()
        self.contact_frames = tape.contact.reshape(tape.frames, tape.num_envs * tape.contact.shape[2], 3).pin_memory()
        self.num_envs, self.frames, self.decimation = tape.num_envs, tape.frames, tape.decimation
        self.frame = 0
        self.sub = 0
        mk = lambda t: [torch.empty_like(t, device=dev) for _ in range(2)]
        self._root_d, self._contact_d = mk(self.root_frames[0]), mk(self.contact_frames[0])
        self._dof_d = [[torch.empty_like(self.dof_frames[0, 0], device=dev) for _ in range(self.decimation)] for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.root_states = self.root_frames[0].to(dev)
        self.dof_state = self.dof_frames[0, 0].to(dev)
        self.contact_forces = self.contact_frames[0].to(dev)
        # a host-side physics needs every sub-step's torques back: staged on the device (double-buffered by frame parity), read back to
        # pinned host memory on a side stream while the next step's inputs arrive (PCIe is full duplex)
        if return_torques:
            nd = self.dof_frames.shape[2] // self.num_envs
            self._tq_stage = [[torch.empty(self.num_envs, nd, device=dev) for _ in range(self.decimation)] for _ in range(2)]
            self.torques_host = [torch.empty(self.num_envs, nd).pin_memory() for _ in range(self.decimation)]
            self._out_stream = torch.cuda.Stream(device=dev)
            self._tq_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._tq_read = [torch.cuda.Event(), torch.cuda.Event()]
        self._prefetch(0)

    def _prefetch(self, frame):
        """H2D of every tensor of `frame` into buffer set frame % 2, on the copy stream."""
        b, f = frame % 2, frame % self.frames
        cs = self._copy_stream
        cs.wait_event(self._consumed[b]) if frame >= 2 else None
        with torch.cuda.stream(cs):
            for j in range(self.decimation):
                self._dof_d[b][j].copy_(self.dof_frames[f, j], non_blocking=True)
            self._root_d[b].copy_(self.root_frames[f], non_blocking=True)
            self._contact_d[b].copy_(self.contact_frames[f], non_blocking=True)
            self._ready[b].record(cs)

    def simulate(self, torques):
        b = self.frame % 2
        if self.sub == 0:
            torch.cuda.current_stream(self.device).wait_event(self._ready[b])
        self.dof_state = self._dof_d[b][self.sub]
        if self.return_torques:
            cur = torch.cuda.current_stream(self.device)
            if self.sub == 0 and self.frame >= 2:
                cur.wait_event(self._tq_read[b])          # the read-back of two steps ago has left this staging set
            self._tq_stage[b][self.sub].copy_(torques, non_blocking=True)
        self.sub += 1

    def refresh(self):
        b = self.frame % 2
        if self.return_torques:
            self._tq_done[b].record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._out_stream):
                self._out_stream.wait_event(self._tq_done[b])
                for j in range(self.decimation):
                    self.torques_host[j].copy_(self._tq_stage[b][j], non_blocking=True)
                self._tq_read[b].record(self._out_stream)
        self.root_states, self.contact_forces = self._root_d[b], self._contact_d[b]
        self.frame += 1
        self.sub = 0
        self._prefetch(self.frame)      # next step's state starts crossing PCIe while this step's post-physics runs

    def commit_resets(self, reset_buf):
        self._consumed[(self.frame - 1) % 2].record(torch.cuda.current_stream(self.device))
