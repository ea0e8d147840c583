// compat: <boost/thread.hpp> (pclvisualizerthread.h:33 in the reference): nothing of it is used without the GUI thread
