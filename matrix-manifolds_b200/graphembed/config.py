"""YAML experiment configs with the reference's `object:` / `closure:` grammar (run.py:138-209,
example_config.yaml): a mapping `{object|closure: {name: dotted.path, params: {...}|[...]}}` is replaced by the
constructed instance, or -- when the constructor still misses positional arguments -- by a callable that takes
them later (`config['embedding'](n_nodes)`, `config['embedding_optimizer'](embedding.xs)`).  Nested nodes are
built innermost first.  Parsed with PyYAML (ruamel.yaml is not required)."""
import importlib

_MARKERS = ('object', 'closure')


def _resolve(name):
    module, _, attr = name.rpartition('.')
    if not module:
        raise ValueError(f'`{name}` is not a dotted path')
    return getattr(importlib.import_module(module), attr)


def _late_bound(ctor, params):
    if isinstance(params, dict):
        return lambda *args, **kwargs: ctor(*args, **kwargs, **params)
    return lambda *args, **kwargs: ctor(*args, *params, **kwargs)


def _instantiate(spec):
    ctor = _resolve(spec['name'])
    params = spec.get('params')
    if params is None:
        params = {}
    try:
        return ctor(**params) if isinstance(params, dict) else ctor(*params)
    except TypeError:  # arguments still missing: bind what we have, take the rest at call time
        return _late_bound(ctor, params)


def build_tree(node):
    """Depth-first construction of every `object`/`closure` node of a parsed YAML tree."""
    if isinstance(node, dict):
        if len(node) == 1:
            (key, value), = node.items()
            if key in _MARKERS and isinstance(value, dict) and 'name' in value:
                return _instantiate({k: build_tree(v) for k, v in value.items()})
        return {k: build_tree(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [build_tree(v) for v in node]
    return node


def parse_config(config_path):
    import yaml
    with open(config_path, 'r') as f:
        return build_tree(yaml.safe_load(f))
