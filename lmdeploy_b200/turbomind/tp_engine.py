"""Control plane of the tensor-parallel engine: one process per GPU, rank 0 owns the request queue and the scheduler
and tells the other ranks what every step contains -- the role the reference gives the TP group's rank 0
(src/turbomind/turbomind.cc:264-266: it owns the request queue; engine/engine.cc broadcasts the scheduled batch).

What travels per step is a few hundred integers (token ids, cumulative lengths, KV block ids); the channel is therefore
a host-side object broadcast over a gloo group, independent of the NVLink data path (lmdeploy_b200.turbomind.tp.TpGroup)
that carries activations.  Every rank keeps a KV pool of identical geometry, so block ids mean the same on all ranks."""
import torch.distributed as dist


class StepChannel:
    def __init__(self, rank, world, group=None):
        """group: a torch.distributed group whose backend can move CPU objects (gloo).  With the default NCCL world of
        a torchrun launch, pass dist.new_group(backend="gloo")."""
        self.rank, self.world, self.group = rank, world, group
        self.open = True

    @property
    def is_leader(self):
        return self.rank == 0

    def publish(self, step_arrays):
        """Leader: send one step's arrays (the dict lmdeploy_b200.turbomind.engine.marshal_step returns)."""
        if not self.is_leader:
            raise RuntimeError("only rank 0 publishes steps")
        if not self.open:
            raise RuntimeError("channel is shut down")
        dist.broadcast_object_list([step_arrays], src=0, group=self.group)

    def receive(self):
        """Follower: block until the leader publishes a step; None = shut down."""
        if self.is_leader:
            raise RuntimeError("rank 0 does not receive steps")
        box = [None]
        dist.broadcast_object_list(box, src=0, group=self.group)
        if box[0] is None:
            self.open = False
        return box[0]

    def shutdown(self):
        """Leader: release the followers (their serve loops return)."""
        if self.is_leader and self.open:
            dist.broadcast_object_list([None], src=0, group=self.group)
            self.open = False


def launch(spec_or_path, engine_config=None, *, seed=0):
    """Run inside every rank of `torchrun --nproc-per-node N` (N = engine_config.tp).  Builds this rank's TP-sharded
    model, the NVLink communicator and the step channel.  Rank 0 gets the engine back (drive it with
    batched_generate / create_instance as usual and call close() when done); every other rank serves steps until the
    leader closes and then returns None."""
    import os

    import torch

    from ..messages import DEFAULT_MAX_BATCH_SIZE, TurbomindEngineConfig
    from .engine import TurboMind
    from .runtime import DecodeModel
    from .specs import SPECS
    from .tp import TpGroup

    cfg = engine_config or TurbomindEngineConfig()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world != cfg.tp:
        raise ValueError(f"engine_config.tp = {cfg.tp} but the launch has {world} ranks")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", str(rank))))
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    control = dist.new_group(backend="gloo")
    eos = None
    if isinstance(spec_or_path, str) and spec_or_path not in SPECS:
        from .loader import load_hf_checkpoint
        spec, model, eos = load_hf_checkpoint(spec_or_path, quant_policy=cfg.quant_policy, block_len=cfg.cache_block_seq_len,
                                              hf_overrides=cfg.hf_overrides, model_format=cfg.model_format,
                                              max_batch=cfg.max_batch_size or DEFAULT_MAX_BATCH_SIZE, tp=world, rank=rank)
    else:
        spec = SPECS[spec_or_path] if isinstance(spec_or_path, str) else spec_or_path
        model = DecodeModel(spec, quant_policy=cfg.quant_policy, block_len=cfg.cache_block_seq_len,
                            max_batch=cfg.max_batch_size or DEFAULT_MAX_BATCH_SIZE, tp_size=world, tp_rank=rank)
        model.load_synthetic(seed)
    tp = TpGroup(model, rank, world)
    channel = StepChannel(rank, world, control)
    # the KV pool must have the same number of blocks on every rank: size it from the smallest free memory
    num_blocks = None
    if cfg.cache_max_entry_count < 1:
        free, _ = torch.cuda.mem_get_info()
        mine = torch.tensor([int(free * cfg.cache_max_entry_count) // (model.layer_bytes * spec.layers)])
        dist.all_reduce(mine, op=dist.ReduceOp.MIN, group=control)
        num_blocks = int(mine.item())
    engine = TurboMind(spec, cfg, model=model, eos_id=eos, num_blocks=num_blocks, channel=channel)
    engine._tp = tp
    if rank == 0:
        return engine
    engine.serve_follower()
    tp.close()
    return None
