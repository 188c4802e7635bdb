class BoxList:  # name only
    pass
