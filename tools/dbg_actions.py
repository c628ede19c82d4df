import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrnco_b200 as rb
from oracle import envs as oenvs, model as omodel, synth
dev = "cuda"
name, n, B = "rcvrp", 12, 2
raw = synth.make_instances(name, B, n, seed=1)
row, col = synth.random_embeddings(B, n + 1, seed=2)
p = omodel.init_decoder_params(name, seed=3)
class Enc(torch.nn.Module):
    def forward(self, td, phase=None):
        return row.to(dev), col.to(dev)
env = rb.RCVRPEnv(generator_params={"num_loc": n}, check_solution=False)
pol = rb.RRNetPolicy(encoder=Enc(), env_name=name).to(dev)
pol.decoder.load_state_dict(p)
for eng in (0, 1):
    rb.set_ffn_engine(eng)
    td = env.reset(rb.TensorDictLite(dict(raw), batch_size=[B]))
    out = pol(td, env, phase="val", decode_type="multistart_greedy", num_starts=13, return_sum_log_likelihood=False)
    print("engine", eng, "actions[:4]\n", out["actions"][:4].cpu(), "\nlogp[:2]\n", out["log_likelihood"][:2].cpu(), "\nreward[:4]", out["reward"][:4].cpu())
