"""keras.optimizers stand-in (TEST INFRASTRUCTURE ONLY): records hyper-parameters."""


class Adam:
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, **kw):
        self.lr, self.beta_1, self.beta_2 = lr, beta_1, beta_2
        self.epsilon = 1e-7 if epsilon is None else epsilon
        self.decay = decay
