import os


def get_original_cwd():
    return os.getcwd()
