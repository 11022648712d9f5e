class MultiControlNetModel:  # name only; golden generation injects a recorded stand-in
    def __init__(self, nets):
        self.nets = list(nets)
