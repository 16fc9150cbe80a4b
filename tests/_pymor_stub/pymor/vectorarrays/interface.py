class VectorArray:
    pass


class VectorSpace:
    pass
