def check_requirements(*a, **k):
    return True


def check_version(*a, **k):
    return True
