import fnmatch


def freeze_modules(model, patterns, recursive=True):
    for name, mod in model.named_modules():
        if any(fnmatch.fnmatchcase(name, p) for p in patterns):
            for prm in mod.parameters(recurse=recursive):
                prm.requires_grad = False
            mod.eval()
    return model
