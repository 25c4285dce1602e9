"""placeholder so that `import matplotlib.pyplot as plt` at the top of the reference modules resolves"""
