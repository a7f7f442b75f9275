"""dm_tree subset used by the reference: map_structure over (nested) dicts / lists / tuples."""


def map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, dict):
        return {k: map_structure(fn, *(s[k] for s in structs)) for k in s0}
    if isinstance(s0, (list, tuple)):
        return type(s0)(map_structure(fn, *items) for items in zip(*structs))
    return fn(*structs)
