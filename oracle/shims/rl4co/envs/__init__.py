from rl4co.envs.common.base import RL4COEnvBase


def get_env(name, *a, **kw):
    raise NotImplementedError("shim: pass an instantiated env")
