function s = read_cfs_fixture(path)
% READ_CFS_FIXTURE  read a flat little-endian "CFSB" fixture (motionplanning_5d_m_b200/fixture_io.py) into a struct.
%   s = read_cfs_fixture('batch_m16ib.bin');   % s.QQ, s.x0 (B x 10), s.ff (B x n), s.caug, s.xref (B x 10H), ...
% Arrays are stored column-major, so reshape() restores them exactly; per-problem data is one ROW per problem
% (transpose to get the n x B columns CFS_FANUC / cfs_mex take in batched use).
fid = fopen(path, 'r', 'ieee-le');
if fid < 0, error('cfs:fixture', 'cannot open %s', path); end
c = onCleanup(@() fclose(fid));
magic = fread(fid, 4, '*char')';
if ~strcmp(magic, 'CFSB'), error('cfs:fixture', 'not a CFSB fixture'); end
version = fread(fid, 1, 'uint32');
if version ~= 1, error('cfs:fixture', 'unsupported version %d', version); end
count = fread(fid, 1, 'uint32');
s = struct();
for k = 1:count
    name = deblank(char(fread(fid, 32, 'uint8')'));
    name = name(name ~= 0);
    ndim = fread(fid, 1, 'uint32');
    dims = fread(fid, ndim, 'uint64')';
    if isempty(dims), dims = [1 1]; end
    if numel(dims) == 1, dims = [dims 1]; end
    s.(name) = reshape(fread(fid, prod(dims), 'double'), dims);
end
end
