class VectorUDT:
    pass
