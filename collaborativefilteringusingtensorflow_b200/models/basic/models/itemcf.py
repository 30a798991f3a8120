"""Item-based neighbourhood model with the reference's constructor and ``train`` entry point (reference
src/models/basic/models/itemcf.py:10-94; driver basic/testicf.py): cosine similarity of the item columns of the training
matrix, the topK neighbours of every item, score(u, j) = sum over the user's items i of sim_K(i, j) * r_ui, masked top-N."""
from ._cf import NeighborhoodModel


class ItemCF(NeighborhoodModel):
    _mode = 'item'
