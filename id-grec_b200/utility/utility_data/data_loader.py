"""utility_data/data_loader.py of the reference, hot-path subset (data_loader.py:8-70,108-133,151-159).

``Data`` keeps the reference's attributes (num_users, num_items, num_nodes, user_item_net,
all_positive, test_dict, train_user, train_item, ...) and adds device-resident copies used by the
CUDA evaluator.  ``sample_data_to_train_all`` reproduces the reference's negative samples and its
numpy global-RNG stream bit for bit, in milliseconds instead of a Python loop over every edge.
"""
import warnings

import numpy as np
import scipy.sparse as sp

warnings.filterwarnings('ignore')


class Data(object):
    def __init__(self, path, config):
        self.path = path
        self.num_users = 0
        self.num_items = 0
        self.num_entities = 0
        self.num_relations = 0
        self.num_nodes = 0
        self.num_train = 0
        self.num_test = 0
        self.config = config
        self._dev = {}
        self.load_data()
        if config:
            self.split_test_dict = None
            self.split_state = None
            if "sparsity_test" in config.keys() and int(config["sparsity_test"]) == 1:
                self.split_test_dict, self.split_state = self.create_sparsity_split()

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, num_users, num_items, train_user, train_item, test_user, test_item, config=None, path="<memory>"):
        """Same object from in-memory COO arrays (synthetic graphs never go through text)."""
        self = cls.__new__(cls)
        self.path, self.config, self._dev = path, config, {}
        self.num_entities = self.num_relations = 0
        self.num_users, self.num_items = int(num_users), int(num_items)
        self.train_user = np.ascontiguousarray(train_user, dtype=np.int64)
        self.train_item = np.ascontiguousarray(train_item, dtype=np.int64)
        self.test_user = np.ascontiguousarray(test_user, dtype=np.int64)
        self.test_item = np.ascontiguousarray(test_item, dtype=np.int64)
        self.num_train, self.num_test = len(self.train_user), len(self.test_user)
        self.pos_length = None
        self.split_test_dict = self.split_state = None
        self._finish()
        if config and int(config.get("sparsity_test", 0)) == 1:
            self.split_test_dict, self.split_state = self.create_sparsity_split()
        return self

    def load_data(self):
        """data_loader.py:27-46."""
        _, self.train_user, self.train_item, self.num_train, self.pos_length = self.read_ratings(self.path + "/train.txt")
        _, self.test_user, self.test_item, self.num_test, _ = self.read_ratings(self.path + "/test.txt")
        self.num_users += 1
        self.num_items += 1
        self.data_statistics_on_load = True
        self._finish()
        self.data_statistics()

    def _finish(self):
        self.num_nodes = self.num_users + self.num_items
        assert len(self.train_user) == len(self.train_item)
        # duplicate (u,i) pairs sum to 2.0 exactly like the reference's csr_matrix call (data_loader.py:42-43)
        self.user_item_net = sp.csr_matrix((np.ones(len(self.train_user)), (self.train_user, self.train_item)),
                                           shape=(self.num_users, self.num_items))
        self.user_item_net.sum_duplicates()
        self.user_item_net.sort_indices()
        self.all_positive = self.get_user_pos_items(range(self.num_users))
        self.test_dict = self.build_test()

    def read_ratings(self, file_name):
        """data_loader.py:48-70: ``user item item ...`` per line; max ids tracked over both files.  One pass over
        the file image by the C entry point idg_parse_ratings instead of a Python loop over the lines."""
        from idgrec import ops
        line_user, line_len, users, items, max_user, max_item = ops.parse_ratings(file_name)
        if len(items):
            self.num_users = max(self.num_users, max_user)
            self.num_items = max(self.num_items, max_item)
        return line_user, users, items, len(items), line_len[line_len > 0].tolist()

    def data_statistics(self):
        print("\t num_users:", self.num_users)
        print("\t num_items:", self.num_items)
        print("\t num_nodes:", self.num_nodes)
        print("\t num_train:", self.num_train)
        print("\t num_test: ", self.num_test)
        print("\t sparisty: ", 1 - (self.num_train + self.num_test) / self.num_users / self.num_items)

    def get_statistics(self):
        strs = "dataset:" + self.config['dataset'] + "\t"
        strs += "num_users:%d, num_items:%d \t" % (self.num_users, self.num_items)
        strs += ("|num_train:%d, num_test:%d, sparsity: %.6f"
                 % (self.num_train, self.num_test, 1 - (self.num_train + self.num_test) / self.num_users / self.num_items))
        return strs

    # -- sampling ----------------------------------------------------------------------------
    def sample_negatives(self):
        """The negatives of data_loader.py:108-127, exact, as one int64 array (edge order = file order).

        The reference draws np.random.randint(0, num_items) once per attempt; scalar draws equal a bulk draw of the
        same length (same values, same final generator state), and a bulk draw equals consecutive smaller ones.  So
        the candidate stream is taken from the global numpy generator in chunks -- the first exactly E long (every edge
        consumes at least one candidate), then short ones for the edges still rejecting -- and replayed against the
        sorted positives of each edge's user by the C entry point idg_neg_sample_walk.  The last chunk is re-drawn at
        exactly the consumed length, so the generator ends where the reference's would (tools.shuffle reads it next)."""
        import ctypes as C
        from idgrec import _lib
        l = _lib.lib()
        E = len(self.train_user)
        neg = np.empty(E, dtype=np.int64)
        if E == 0:
            return neg
        if getattr(self, "_walk_csr", None) is None:
            net = self.user_item_net
            self._walk_csr = (np.ascontiguousarray(self.train_user, dtype=np.int64), np.ascontiguousarray(net.indptr, dtype=np.int32),
                              np.ascontiguousarray(net.indices, dtype=np.int32))
        tu, ip, ix = self._walk_csr
        e, chunk = 0, E
        done, used = C.c_int64(0), C.c_int64(0)
        while e < E:
            state = np.random.get_state()
            cand = np.random.randint(0, self.num_items, size=chunk)
            if cand.dtype != np.int64:
                cand = cand.astype(np.int64)
            _lib.check(l.idg_neg_sample_walk(tu.ctypes.data, e, E, ip.ctypes.data, ix.ctypes.data, cand.ctypes.data, len(cand),
                                             neg.ctypes.data, C.byref(done), C.byref(used)), "idg_neg_sample_walk")
            e = int(done.value)
            if e >= E and int(used.value) < len(cand):
                np.random.set_state(state)                      # finished inside this chunk: consume exactly `used` draws
                np.random.randint(0, self.num_items, size=int(used.value))
            chunk = max(1024, 2 * (E - e))
        return neg

    def sample_data_to_train_all(self):
        """data_loader.py:108-127 -> int64 [E, 3] (user, positive, negative); see sample_negatives."""
        if len(self.train_user) == 0:
            return np.zeros((0, 3), dtype=np.int64)
        return np.stack([self.train_user, self.train_item, self.sample_negatives()], axis=1)

    def sample_data_to_train_random(self):
        """data_loader.py:86-106 (uniform user sampling of the official LightGCN code; no model of the reference calls
        it).  The same numpy calls in the same order, so the global generator stream and the result are the reference's."""
        users = np.random.randint(0, self.num_users, len(self.train_user))
        sample_list = []
        ip, ix = self.user_item_net.indptr, self.user_item_net.indices
        for user in users:
            lo, hi = ip[user], ip[user + 1]
            if hi == lo:
                continue
            positive_item = ix[lo + np.random.randint(0, hi - lo)]
            while True:
                negative_item = np.random.randint(0, self.num_items)
                k = lo + np.searchsorted(ix[lo:hi], negative_item)
                if k < hi and ix[k] == negative_item:
                    continue
                break
            sample_list.append([user, positive_item, negative_item])
        return np.array(sample_list)

    def get_user_n_neg_items(self, users, n):
        """data_loader.py:135-149: n negatives per user, rejection against the user's train positives."""
        ip, ix = self.user_item_net.indptr, self.user_item_net.indices
        negative_items = []
        for user in users:
            lo, hi = ip[user], ip[user + 1]
            negative_list = []
            for _ in range(n):
                while True:
                    negative_item = np.random.randint(0, self.num_items)
                    k = lo + np.searchsorted(ix[lo:hi], negative_item)
                    if k < hi and ix[k] == negative_item:
                        continue
                    negative_list.append(negative_item)
                    break
            negative_items.append(negative_list)
        return negative_items

    def get_user_pos_items(self, users):
        """data_loader.py:129-133: sorted train items of each user (views into the CSR)."""
        ip, ix = self.user_item_net.indptr, self.user_item_net.indices
        return [ix[ip[u]:ip[u + 1]] for u in users]

    def build_test(self):
        """data_loader.py:151-159: {user: [items in file order]} for users with a non-empty line."""
        test_data = {}
        if len(self.test_user) == 0:
            return test_data
        order = np.argsort(self.test_user, kind="stable")
        su, si = self.test_user[order], self.test_item[order]
        starts = np.flatnonzero(np.concatenate([[True], su[1:] != su[:-1]]))
        ends = np.concatenate([starts[1:], [len(su)]])
        first_seen = {}
        for s, e in zip(starts, ends):
            first_seen[int(su[s])] = (int(order[s]), si[s:e].tolist())
        for u, (_, items) in sorted(first_seen.items(), key=lambda kv: kv[1][0]):
            test_data[u] = items
        return test_data

    def create_sparsity_split(self):
        """data_loader.py:161-204: test users bucketed by activity (train + test interactions), cut
        every time a bucket sequence accumulates a quarter of all interactions (the reference never
        advances its ``count``, so the threshold stays at 25 %), remainder appended at the end (an
        EMPTY remainder is appended too when the last activity level itself closes a group -- kept,
        see batch_test.sparsity_test).  Users inside a group keep test_dict order within each level."""
        users = np.fromiter(self.test_dict.keys(), dtype=np.int64, count=len(self.test_dict))
        ip = self.user_item_net.indptr
        n_test = np.fromiter((len(v) for v in self.test_dict.values()), dtype=np.int64, count=len(users))
        activity = (ip[users + 1] - ip[users]).astype(np.int64) + n_test
        order = np.argsort(activity, kind="stable")          # level ascending, test_dict order inside a level
        levels, first = np.unique(activity[order], return_index=True)
        bounds = np.append(first, len(order))
        total = self.num_train + self.num_test
        split_uids, split_state = [], []
        temp, n_rates, n_count = [], 0, total

        def close(level):
            split_uids.append(temp)
            state = '\t #inter per user<=[%d], #users=[%d], #all rates=[%d]' % (level, len(temp), n_rates)
            split_state.append(state)
            print(state)

        for idx, level in enumerate(levels.tolist()):
            members = users[order[bounds[idx]:bounds[idx + 1]]].tolist()
            temp = temp + members
            n_rates += level * len(members)
            n_count -= level * len(members)
            if n_rates >= 0.25 * total:
                close(level)
                temp, n_rates = [], 0
            if idx == len(levels) - 1 or n_count == 0:
                close(level)
        return split_uids, split_state

    # -- device-side copies for the CUDA evaluator ------------------------------------------
    def device_cache(self, device):
        """mask CSR (train positives), test CSR (sorted, duplicates kept) and the test-user list."""
        import torch
        key = str(device)
        if key not in self._dev:
            net = self.user_item_net
            users = np.fromiter(self.test_dict.keys(), dtype=np.int64, count=len(self.test_dict))
            tu, ti = self.test_user, self.test_item
            order = np.lexsort((ti, tu))
            tptr = np.zeros(self.num_users + 1, dtype=np.int64)
            np.add.at(tptr, tu + 1, 1)
            tptr = np.cumsum(tptr)
            self._dev[key] = {
                "mask_indptr": torch.from_numpy(net.indptr.astype(np.int32)).to(device),
                "mask_indices": torch.from_numpy(net.indices.astype(np.int32)).to(device),
                "test_indptr": torch.from_numpy(tptr.astype(np.int32)).to(device),
                "test_indices": torch.from_numpy(ti[order].astype(np.int32)).to(device),
                "test_users": torch.from_numpy(users).to(device),
                "train_user": torch.from_numpy(np.ascontiguousarray(self.train_user, dtype=np.int64)).to(device),
                "train_item": torch.from_numpy(np.ascontiguousarray(self.train_item, dtype=np.int64)).to(device),
            }
        return self._dev[key]
