class LayerRecipe:
    pass
