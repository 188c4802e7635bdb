class BoxCoder:  # name only (model/inference.py:3 imports it, never uses it)
    pass
