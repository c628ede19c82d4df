from rl4co.models.common.constructive import AutoregressiveDecoder  # noqa: F401
