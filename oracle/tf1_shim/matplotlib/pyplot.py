def _absent(*_a, **_k):
    raise NotImplementedError("matplotlib is not installed; plotting is outside the paths the fixtures exercise")


subplots = figure = savefig = close = imshow = colorbar = xlabel = ylabel = title = tight_layout = bar = xticks = _absent
