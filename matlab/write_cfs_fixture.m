function write_cfs_fixture(path, s)
% WRITE_CFS_FIXTURE  write the fields of struct s (real double arrays) as a flat little-endian "CFSB" fixture, the format of
% motionplanning_5d_m_b200/fixture_io.py (read back by fixture_io.read_fixture and by read_cfs_fixture.m).
names = fieldnames(s);
fid = fopen(path, 'w', 'ieee-le');
if fid < 0, error('cfs:fixture', 'cannot open %s for writing', path); end
c = onCleanup(@() fclose(fid));
fwrite(fid, 'CFSB', 'char');
fwrite(fid, 1, 'uint32');
fwrite(fid, numel(names), 'uint32');
for k = 1:numel(names)
    name = names{k};
    if numel(name) > 31, error('cfs:fixture', 'array name too long: %s', name); end
    a = double(s.(name));
    fwrite(fid, [uint8(name) zeros(1, 32 - numel(name), 'uint8')], 'uint8');
    dims = size(a);
    fwrite(fid, numel(dims), 'uint32');
    fwrite(fid, dims, 'uint64');
    fwrite(fid, a(:), 'double');          % column-major, as stored
end
end
