FrozenVariableDict = dict
