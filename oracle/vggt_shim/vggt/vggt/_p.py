"""Helper: local parameter dict of a module (names relative to the module)."""


def params_of(module):
    d = dict(module.named_parameters())
    d.update(dict(module.named_buffers()))
    return d
