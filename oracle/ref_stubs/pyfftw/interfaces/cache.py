def enable():
    return None
