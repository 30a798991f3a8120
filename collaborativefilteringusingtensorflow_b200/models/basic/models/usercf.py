"""User-based neighbourhood model with the reference's constructor and ``train`` entry point (reference
src/models/basic/models/usercf.py:10-90; driver basic/testucf.py): cosine similarity of the user rows, score(u, j) = sum over
the topK most similar users v (positive similarity only) of sim(u, v) * r_vj, masked top-N."""
from ._cf import NeighborhoodModel


class UserCF(NeighborhoodModel):
    _mode = 'user'
