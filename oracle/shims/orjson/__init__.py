import json


def loads(s):
    return json.loads(s)


def dumps(o):
    return json.dumps(o).encode()
